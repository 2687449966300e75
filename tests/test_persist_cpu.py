"""CPU tests of the persistence registry and of the checkpoint file format (aces4_b200/csrc/persist.cu) against a
struct-level restatement of the reference's stream format (worker_persistent_array_manager.cpp:157-260,
setup/io_utils.cpp:44-70,114-155).  Only scalars can be exercised without a GPU; arrays are tests/test_gpu_persist.py."""
import struct

import pytest


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.build()
    return s.api


def ref_string(s):
    s = s.rstrip(" ")
    return struct.pack("<i", len(s) + 1) + s.encode() + b"\0"


def ref_checkpoint(scalars, arrays=()):
    """bytes the reference's checkpoint_persistent writes (arrays: (label, dims<=6, flat column-major data))"""
    out = struct.pack("<i", len(scalars))
    for k in sorted(scalars):   # std::map iteration order
        out += ref_string(k) + struct.pack("<d", scalars[k])
    out += struct.pack("<i", len(arrays))
    for label, dims, data in sorted(arrays):
        dims6 = list(dims) + [1] * (6 - len(dims))
        out += ref_string(label) + struct.pack("<i", 6) + struct.pack("<i6i", 6, *dims6)
        out += struct.pack("<i", len(data)) + struct.pack(f"<{len(data)}d", *data)
    return out


def test_scalar_registry_semantics(sip):
    sip.persist_scalar("e_scf", -75.58432674274046)
    sip.persist_scalar("e_scf", -75.5)            # a repeated label overwrites
    sip.persist_scalar("ecorr", -0.12610179886435)
    assert sip.persist_counts()[0] == 2
    assert sip.restore_scalar("e_scf") == -75.5
    with pytest.raises(sip.SipGpuError):          # restore erases the entry
        sip.restore_scalar("e_scf")
    assert sip.restore_scalar("ecorr") == -0.12610179886435
    assert sip.persist_counts() == (0, 0, 0)


def test_checkpoint_bytes_match_the_reference_format(sip, tmp_path):
    scal = {"scf_energy": -75.58432674274046, "lccd_correlation": -0.12610179886435, "padded label  ": 3.5}
    for k, v in scal.items():
        sip.persist_scalar(k, v)
    path = tmp_path / "worker.ckpt"
    sip.persist_checkpoint(path)
    # trailing blanks are trimmed by write_string; std::map orders by the untrimmed key
    want = struct.pack("<i", 3)
    for k in sorted(scal):
        want += ref_string(k) + struct.pack("<d", scal[k])
    want += struct.pack("<i", 0)
    assert path.read_bytes() == want
    # init_from_checkpoint refuses a non-empty registry, then restores from the file
    with pytest.raises(sip.SipGpuError):
        sip.persist_init_from_checkpoint(path)
    for k in scal:
        sip.restore_scalar(k)
    sip.persist_init_from_checkpoint(path)
    assert sip.restore_scalar("padded label") == 3.5
    assert sip.restore_scalar("scf_energy") == -75.58432674274046
    assert sip.restore_scalar("lccd_correlation") == -0.12610179886435


def test_reference_written_checkpoint_restores(sip, tmp_path):
    path = tmp_path / "ref.ckpt"
    path.write_bytes(ref_checkpoint({"a": 1.25, "b": -2.0}))
    sip.persist_init_from_checkpoint(path)
    assert sip.restore_scalar("a") == 1.25 and sip.restore_scalar("b") == -2.0


def test_malformed_checkpoint_is_an_error(sip, tmp_path):
    path = tmp_path / "bad.ckpt"
    path.write_bytes(ref_checkpoint({"a": 1.25})[:-7])
    with pytest.raises(sip.SipGpuError):
        sip.persist_init_from_checkpoint(path)
    sip.lib().sipgpu_restore_scalar(b"a", None)
    with pytest.raises(sip.SipGpuError):
        sip.persist_init_from_checkpoint(tmp_path / "missing.ckpt")
    # drain whatever the truncated file left behind
    n = sip.persist_counts()[0]
    assert n in (0, 1)
    if n:
        sip.restore_scalar("a")
