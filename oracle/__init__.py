"""CPU oracle for the TinyDiffusionModels hot path — TEST INFRASTRUCTURE ONLY.

Nothing under oracle/ may be imported by the product package (tinydiffusionmodels_b200/, src/).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
and only as the checker or the CPU baseline, never as the thing shipped.
"""
