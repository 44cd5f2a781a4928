"""CPU oracle package — test infrastructure only (see gp_oracle.py / gp_oracle.c headers)."""
