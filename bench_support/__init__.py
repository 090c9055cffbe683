"""Synthetic-input generators shared by tests/ and bench.py (not product code)."""
