"""CPU oracle for the metran Kalman path -- TEST INFRASTRUCTURE, not product.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
