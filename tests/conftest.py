import os
import sys

import pytest

# tests/test_shard_group_gpu.py runs two ranks of a model-parallel group on ONE device (two streams): a rank's flag-wait
# kernel spins until the other rank's kernels have published.  With CUDA's default LAZY module loading the first launch
# of any kernel (ours, CUB's, the driver's own memset kernels) stalls until the kernels already running on the device
# finish — here the spinner, which waits for that very launch (measured: tools/stream_alias_probe.cu, 1 of 552 stream
# pairs serialised with lazy loading, 0 with eager).  Must be set before the CUDA context exists.  Deployments with one
# rank per device are not affected: a module load only waits for kernels of its own device.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.join(ROOT, "tests")
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)  # `import fake_triton` (the harness package under tests/)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _native_built() -> bool:
    return os.path.exists(os.path.join(ROOT, "hugectr_backend_b200", "lib", "libhpsx.so"))


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Build the native libraries once per test session when they are missing (CPU box: nvcc cross-compiles)."""
    if not _native_built():
        import __graft_entry__ as g

        g.build()
    from oracle import hps_oracle

    hps_oracle.build_c_oracle()
    yield


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible — there is no CPU fallback for the GPU path")
    torch.cuda.set_device(0)
    return 0
