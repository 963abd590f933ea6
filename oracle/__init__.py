"""CPU oracle — test infrastructure only (see mr_oracle.c)."""
