"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see gcm_oracle.c header)."""
