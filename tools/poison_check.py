"""Race / stale-read evidence without compute-sanitizer (closed on the pool): build libcudampc_poison.so with -DMPC_POISON - the
right-hand side / solution rows are overwritten with NaN at the point of every ADMM iteration where they are dead - and run the
GPU parity tests against it (CUDAMPC_LIB).  A phase that consumed a row before its producer wrote it (missing barrier between the
two warps of a group at horizons >= 85, wrong row index) would read the NaN.
    python tools/poison_check.py --build          # here (nvcc cross-compiles)
    python tools/poison_check.py                  # on the GPU box: runs pytest with the poisoned library
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = os.path.join(ROOT, "rrt_mpc_b200", "libcudampc_poison.so")
if "--build" in sys.argv or not os.path.exists(lib):
    import __graft_entry__ as G
    G.build_cuda(lib, extra=["-DMPC_POISON"], tag="_poison")
    if "--build" in sys.argv:
        sys.exit(0)
env = dict(os.environ, CUDAMPC_LIB=lib)
sys.exit(subprocess.call([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_gpu_rollout.py", "-q", "-x", "-k", "not extension_is_loaded"], cwd=ROOT, env=env))
