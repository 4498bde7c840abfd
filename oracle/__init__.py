"""CPU oracle of the guardX env step.  TEST INFRASTRUCTURE ONLY (see gx_oracle.c)."""
