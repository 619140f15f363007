"""CPU oracle (test infrastructure only) -- see swrt_oracle.py / swrt_oracle.c headers."""
