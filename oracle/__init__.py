"""CPU oracle for the PFEM3D hot path -- TEST INFRASTRUCTURE ONLY (see pfem_oracle.cpp header)."""
