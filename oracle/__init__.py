"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/ar_oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  The product never does.
"""
