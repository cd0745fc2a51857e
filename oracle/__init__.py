"""CPU oracles for the forced-alignment decoder -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; hubertfa_b200/ never does (tests/test_host_logic.py::test_product_never_touches_the_oracle enforces it).
"""
