"""Argument contract of the out-of-core feed (mininf_b200/stream.py); the staged copies themselves
are covered on the GPU in tests/test_engine_gpu.py."""
import pytest
import torch

from mininf_b200.stream import HostBatchStream


def test_host_batch_stream_rejects_bad_arguments():
    X, y = torch.zeros(10, 3), torch.zeros(10)
    with pytest.raises(ValueError, match="at least one tensor"):
        HostBatchStream({}, 4)
    with pytest.raises(ValueError, match="leading dimension"):
        HostBatchStream({"X": X, "y": y[:9]}, 4)
    with pytest.raises(ValueError, match="batch_rows must be positive"):
        HostBatchStream({"X": X, "y": y}, 0)
    with pytest.raises(ValueError, match="depth at least 2"):
        HostBatchStream({"X": X, "y": y}, 4, depth=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        HostBatchStream({"X": X, "y": y}, 4, device="cpu")
