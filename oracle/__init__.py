"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement + reference loader).

Import rules (enforced by tests/test_no_oracle_in_product.py): only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
