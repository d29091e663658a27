"""import-time placeholder (utilities.py:16)"""
