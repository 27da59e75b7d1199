"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithm for the YOLO11 inference hot path, used as the checker by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under
yolo-lite_b200/ may import this package.
"""
