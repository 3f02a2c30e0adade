"""CPU oracle of the path-tracing hot path — TEST INFRASTRUCTURE (see oracle/oracle.h).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
