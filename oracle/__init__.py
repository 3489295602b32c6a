"""CPU oracle of the PGBART step — test infrastructure only (see pgbart_oracle.c)."""
