"""CPU oracle: test infrastructure only (see torch_oracle.py / gs_oracle.c headers)."""
