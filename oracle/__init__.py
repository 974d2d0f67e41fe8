"""CPU oracle (test infrastructure only) — see hdrtvnet_oracle.py."""
