"""CPU oracle — test infrastructure only (see oracle/oracle.c and oracle/oracle.py)."""
