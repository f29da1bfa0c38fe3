"""Test harness only: the Python daemon (cmd/nvidia.py -> gpumanager.Run) on a GPU-less box. It swaps in the synthetic
inventory of tests/fakes.py for the calls that need a driver (init/shutdown, enumeration, health-thread start) and
then runs the unmodified entry point, so the manager loop, the watchers, the dump and the exit codes are real."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gpushare_device_plugin_b200 import device  # noqa: E402
from gpushare_device_plugin_b200.nvidia import nvidia  # noqa: E402
from tests import fakes  # noqa: E402


class _Patch:
    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


n_gpus = int(os.environ.get("PY_DAEMON_FAKE_GPUS", "8"))
fakes.install(_Patch, n_gpus=n_gpus)
device.init = lambda: None
device.shutdown = lambda: None
nvidia.getDeviceCount = lambda: n_gpus

from gpushare_device_plugin_b200.cmd import nvidia as cmd  # noqa: E402

cmd.main(sys.argv[1:])
