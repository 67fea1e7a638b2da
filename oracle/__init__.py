"""CPU oracle for the jss-v1 hot path -- TEST INFRASTRUCTURE ONLY (see jss_oracle.c)."""
