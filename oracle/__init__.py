"""CPU oracle for the nav24 ORB front end — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (nav24_b200/) never does.
"""
