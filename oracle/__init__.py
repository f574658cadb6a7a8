"""CPU oracles: TEST INFRASTRUCTURE ONLY (see each module's header).
Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU legs."""
