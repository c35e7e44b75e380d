"""Parity oracle (test infrastructure only -- see pmf_oracle.py header)."""
