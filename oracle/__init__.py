"""CPU oracle for libsphb200 — TEST INFRASTRUCTURE ONLY (see sph_oracle.cpp)."""
