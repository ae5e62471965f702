"""TEST INFRASTRUCTURE -- CPU oracle for the cluster-expansion Metropolis path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (cemc_b200/) never does.
"""
