"""ORACLE package: CPU restatements of the reference hot path, used ONLY as a checker by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product package never
imports anything from here (tests/test_no_oracle_in_product.py enforces it)."""
