"""import-time placeholder (utilities.py:14)"""
