"""CPU oracle of the reference's evaluation path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may
import this package (see oracle/np_oracle.py).  The product package pychebyshev_b200 never does.
"""
