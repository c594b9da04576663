import sys, runpy
sys.argv = ['bench.py', '--steps', '3', '--warmup', '3', '--e2e-steps', '1', '--no-extras']
import oracle
def _skip():
    raise RuntimeError("cpu baseline skipped in this quick check")
oracle.Ref = _skip
runpy.run_path('bench.py', run_name='__main__')
