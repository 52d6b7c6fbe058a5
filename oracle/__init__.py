"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.c, oracle/pyref.py).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  nova_b200/ must never import it.
"""
