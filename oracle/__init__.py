"""TEST INFRASTRUCTURE: CPU oracle for the BFM reconstruction + rasterization path.

Nothing under voicepuppet_b200/ imports this package.  Allowed users: tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
