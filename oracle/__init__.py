"""oracle/ — CPU restatements of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never from the
product package.  See DESIGN.md §Oracle.
"""
