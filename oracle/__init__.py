"""CPU oracle (test infrastructure only -- see oracle/bpr_oracle.py)."""
