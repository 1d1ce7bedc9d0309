"""CPU oracle of the scheduler hot path — TEST INFRASTRUCTURE ONLY (see oracle/lig_oracle.h)."""
