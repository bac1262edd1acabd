"""Input generators for tests and bench.py (NOT part of the product package): real designs of the
named circuits restated from the reference's design generator (circuit_design.py)."""
