"""CPU oracle (test infrastructure only; see oracle/shapes_oracle.h).

Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports it.
"""
