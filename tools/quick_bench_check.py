"""Scratch: bench.py's GPU arm (3 steps, no extras) with the ~15 s CPU-baseline leg stubbed out -- a quick check that the
one JSON line still comes out after an edit of bench.py.  Not a measurement."""
import sys, runpy
sys.argv = ['bench.py', '--steps', '3', '--warmup', '3', '--e2e-steps', '1', '--no-extras']
import oracle
def _skip():
    raise RuntimeError("cpu baseline skipped in this quick check")
oracle.Ref = _skip
runpy.run_path('bench.py', run_name='__main__')
