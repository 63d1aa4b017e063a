"""Oracle package: CPU restatements of the reference lattice-settle path.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and the
CPU-baseline legs of bench.py; the product package `oscillink_b200` must never import
it (tests/test_no_oracle_in_product.py enforces that).
"""
