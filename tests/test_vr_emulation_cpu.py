"""The two variational-refinement kernels of csrc/dis.cu compiled for the host (tests/emul: the kernel SOURCE between
the //@emul markers, one host thread per CUDA thread, a pthread barrier for __syncthreads / barrier.cluster, arrays for
distributed shared memory): the shared-memory / register resident kernel must leave the same bits as the streaming
cluster kernel -- whose GPU output is pinned bit-exact to cv2 by tests/test_dis_gpu.py -- on single-CTA levels, row
bands with halo pushes, odd sizes and empty bands.  Catches indexing and synchronisation mistakes without a GPU."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("vr_emul")
    subprocess.run(["sh", os.path.join(HERE, "emul", "build.sh"), str(out)], check=True, capture_output=True, text=True)
    return os.path.join(str(out), "vr_emul")


CASES = [
    (120, 67, 8, 1, 1024),   # level 3 of a 960x540 working image: one CTA per pair
    (120, 67, 8, 1, 512),
    (240, 135, 8, 4, 1024),  # its finest level: four row bands, halo rows pushed through (emulated) distributed shared memory
    (101, 57, 4, 2, 512),    # odd sizes, two bands
    (50, 33, 2, 8, 512),     # more bands than rows need: thin and empty bands
    (60, 33, 1, 1, 1024),    # the two coarsest levels: fewer cells than threads, empty cell slots are skipped
    (30, 16, 1, 1, 1024),
    (16, 200, 4, 2, 512),    # tall and narrow: 100 rows per band (the packed cell word has 7 bits for the row)
]


@pytest.mark.parametrize("case", CASES, ids=["x".join(map(str, c)) for c in CASES])
def test_resident_kernel_equals_cluster_kernel(emulator, case):
    run = subprocess.run([emulator] + [str(v) for v in case], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "0 of" in run.stdout and "FAIL" not in run.stdout
