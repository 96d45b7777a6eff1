"""CPU oracle for the gan-ode latent-motion hot path.  TEST INFRASTRUCTURE ONLY: see the header of
oracle/torchdiffeq_restatement.py for who may import this package (tests/, smoke(), bench.py's CPU legs)."""
