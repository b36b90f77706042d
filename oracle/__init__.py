"""TEST INFRASTRUCTURE: CPU restatement of the reference's hot path (the parity oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package, and only as the checker or the timed CPU baseline.  The product
package (eyediseasesegmentation_b200/) never imports it and has no CPU fallback.
"""
