"""CPU oracle (test infrastructure only; see s4f_oracle.cpp)."""
