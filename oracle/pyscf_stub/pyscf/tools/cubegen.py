"""import-time placeholder (utilities.py:15)"""
