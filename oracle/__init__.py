"""CPU oracle for the nway hot path -- test infrastructure only (see nway_oracle.py)."""
