"""CPU oracle for the haloop alignment-loss hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package. haloop_b200/ (the product) never does.
"""
