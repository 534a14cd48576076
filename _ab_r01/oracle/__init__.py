"""CPU oracle of the reference D2Q9 MRT-LES time step -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from the product package.
"""
