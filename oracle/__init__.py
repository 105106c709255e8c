"""CPU oracle for the BoneJ Thickness hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  Nothing under bonej2_b200/ does.
"""
