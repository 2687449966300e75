"""GPU tests of persistence: label hand-off of resident arrays between SIAL programs, the worker
checkpoint file in the reference's byte format, and the per-rank array files."""
import struct

import numpy as np
import pytest

from test_persist_cpu import ref_checkpoint

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.init(0)
    return s.api


def test_contiguous_handoff_keeps_the_block_resident(sip):
    a = np.asfortranarray(np.arange(60.0).reshape(3, 4, 5))
    blk = sip.DeviceBlock.from_numpy(a)
    ptr = blk.ptr
    sip.persist_contiguous("ca", blk)
    assert sip.persist_counts()[1] == 1
    back = sip.restore_contiguous("ca")
    assert back.ptr == ptr and back.shape == (3, 4, 5)      # ownership transfer, no copy
    assert np.array_equal(back.to_numpy(), a)
    with pytest.raises(sip.SipGpuError):
        sip.restore_contiguous("ca")


def test_checkpoint_with_arrays_round_trip_and_reference_bytes(sip, tmp_path):
    rng = np.random.default_rng(3)
    fock = np.asfortranarray(rng.uniform(-1, 1, (13, 13)))
    ca = np.asfortranarray(rng.uniform(-1, 1, (13, 5)))
    sip.persist_scalar("scf_energy", -75.58432674274046)
    sip.persist_contiguous("fock_a", sip.DeviceBlock.from_numpy(fock))
    sip.persist_contiguous("ca", sip.DeviceBlock.from_numpy(ca))
    path = tmp_path / "worker.ckpt"
    sip.persist_checkpoint(path)
    want = ref_checkpoint({"scf_energy": -75.58432674274046},
                          [("fock_a", fock.shape, fock.ravel(order="F").tolist()), ("ca", ca.shape, ca.ravel(order="F").tolist())])
    assert path.read_bytes() == want
    for label in ("fock_a", "ca"):
        sip.restore_contiguous(label).free()
    sip.restore_scalar("scf_energy")
    sip.persist_init_from_checkpoint(path)                 # a restart
    got = sip.restore_contiguous("fock_a")
    assert got.shape == (13, 13, 1, 1, 1, 1) and np.array_equal(got.to_numpy().reshape(13, 13, order="F"), fock)
    got = sip.restore_contiguous("ca")
    assert np.array_equal(got.to_numpy().reshape(13, 5, order="F"), ca)
    assert sip.restore_scalar("scf_energy") == -75.58432674274046


def test_distributed_array_label_handoff_and_files(sip, tmp_path):
    segs = [[3, 4], [2, 2, 3], [3, 4], [2, 2, 3]]
    A = sip.DistArray(segs)
    rng = np.random.default_rng(9)
    blocks = {}
    for idx in np.ndindex(2, 3, 2, 3):
        idx1 = tuple(i + 1 for i in idx)
        b = np.asfortranarray(rng.uniform(-1, 1, A.block_shape(idx1)))
        blocks[idx1] = b
        A.put(idx1, sip.DeviceBlock.from_numpy(b))
    sip.sync()
    base = A.local_base()
    A.save(tmp_path / "job.T2.1.parr", tmp_path / "job.T2.1.parr_index")
    # file structure: <int chunk_size><int servers><doubles>, index <77><nblocks><offsets>
    raw = (tmp_path / "job.T2.1.parr").read_bytes()
    chunk, servers = struct.unpack("<ii", raw[:8])
    assert servers == 1 and chunk * 8 == len(raw) - 8 == A.local_bytes()
    index = np.frombuffer((tmp_path / "job.T2.1.parr_index").read_bytes(), dtype="<i8")
    assert index[0] == 77 and index[1] == 36 and len(index) == 38
    for idx1, b in blocks.items():
        off = int(index[2 + A.block_number(idx1)])
        got = np.frombuffer(raw[off: off + 8 * b.size], dtype="<f8")
        assert np.array_equal(got, b.ravel(order="F"))
    # set_persistent / restore_persistent: the slab is adopted by the next program's array, no copy
    A.persist("T2_amplitudes")
    B = sip.DistArray(segs)
    B.restore("T2_amplitudes")
    assert B.local_base() == base
    for idx1, b in blocks.items():
        assert np.array_equal(B.get(idx1).to_numpy(), b)
    # restart from the files into a fresh array; a different layout is rejected
    Cc = sip.DistArray(segs)
    Cc.load(tmp_path / "job.T2.1.parr", tmp_path / "job.T2.1.parr_index")
    for idx1, b in blocks.items():
        assert np.array_equal(Cc.get(idx1).to_numpy(), b)
    Dd = sip.DistArray([[3, 4], [2, 2, 3], [4, 3], [2, 2, 3]])
    with pytest.raises(sip.SipGpuError):
        Dd.load(tmp_path / "job.T2.1.parr", tmp_path / "job.T2.1.parr_index")
    for X in (B, Cc, Dd):
        X.destroy()


def test_access_tracking_feeds_the_race_detector(sip):
    """get / put / put += record what this rank touched between barriers (dist.cu); the summary goes through
    sipgpu_consistency_validate together with a simulated second worker's summary."""
    A = sip.DistArray([[2, 3], [2, 3]])
    A.track_accesses(True)
    blk = sip.DeviceBlock(A.block_shape((1, 2))).fill(1.0)
    A.put_accumulate((1, 2), blk)
    A.put_accumulate((1, 2), blk)
    g = A.get((2, 1))
    A.put((2, 2), sip.DeviceBlock(A.block_shape((2, 2))).fill(3.0))
    mine = A.section_accesses()
    n12, n21, n22 = A.block_number((1, 2)), A.block_number((2, 1)), A.block_number((2, 2))
    assert sorted(mine) == sorted([(n12, sip.ACCESS_PUT_ACCUMULATE), (n21, sip.ACCESS_GET), (n22, sip.ACCESS_PUT)])
    other_ok = [(n12, sip.ACCESS_PUT_ACCUMULATE, 1), (n21, sip.ACCESS_GET, 1)]
    sip.consistency_validate([(b, f, 0) for b, f in mine] + other_ok)
    with pytest.raises(sip.SipGpuError):
        sip.consistency_validate([(b, f, 0) for b, f in mine] + [(n22, sip.ACCESS_GET, 1)])
    A.section_reset()
    assert A.section_accesses() == []
    assert np.all(A.get((1, 2)).to_numpy() == 2.0) and g.shape == A.block_shape((2, 1))
    A.destroy()


def test_mirrored_block_coherence(sip):
    """BlockManager::lazy_gpu_* semantics on real buffers: copies happen only when the side being used is missing or
    stale, and the data seen on either side is always the latest write."""
    a = np.asfortranarray(np.arange(24.0).reshape(4, 6))
    m = sip.MirroredBlock(a.copy(order="F"))               # the mirror writes back into ITS host buffer, not into `a`
    assert m.status() == sip.ON_HOST
    d = m.on_device(sip.READ_ON_DEVICE)                     # allocate + h2d
    assert m.status() == sip.ON_HOST | sip.ON_GPU and np.array_equal(d.to_numpy(), a)
    d = m.on_device(sip.UPDATE_ON_DEVICE)
    d.scale(2.0)                                            # the device copy is now the newer one
    assert m.status() & sip.DIRTY_ON_GPU
    h = m.on_host(sip.READ_ON_HOST)                         # d2h because dirty on gpu
    assert np.array_equal(h, 2.0 * a) and not (m.status() & sip.DIRTY_ON_GPU)
    h = m.on_host(sip.WRITE_ON_HOST)
    h[...] = 7.0                                            # host is newer
    assert m.status() & sip.DIRTY_ON_HOST
    d = m.on_device(sip.READ_ON_DEVICE)                     # h2d because dirty on host
    assert np.all(d.to_numpy() == 7.0) and not (m.status() & sip.DIRTY_ON_HOST)
    m.destroy()
    # a block that exists nowhere: write_on_device creates it (zeroed) on the device, read_on_host brings it over
    m2 = sip.MirroredBlock(shape=(3, 5))
    with pytest.raises(sip.SipGpuError):
        m2.on_device(sip.READ_ON_DEVICE)                    # "block allocated neither on host or gpu"
    d = m2.on_device(sip.WRITE_ON_DEVICE)
    d.increment(1.5)
    assert m2.status() == sip.ON_GPU | sip.DIRTY_ON_GPU
    assert np.all(m2.on_host(sip.READ_ON_HOST) == 1.5) and m2.status() == sip.ON_GPU | sip.ON_HOST
    m2.destroy()


def test_mirrored_block_inside_a_recording(sip):
    """sipgpu.h: blocking calls flush an open recording.  A contraction RECORDED into a mirror's device side must be
    visible to a host access made inside the same recording (the d2h may not overtake the recorded writer), a copy towards
    the device may not overtake recorded readers of the old contents, and destroying the mirror defers the free."""
    rng = np.random.default_rng(5)
    L = np.asfortranarray(rng.uniform(-1, 1, (6, 5)))
    R = np.asfortranarray(rng.uniform(-1, 1, (5, 7)))
    dL, dR = sip.DeviceBlock.from_numpy(L), sip.DeviceBlock.from_numpy(R)
    m = sip.MirroredBlock(np.zeros((6, 7), order="F"))
    old = sip.MirroredBlock(np.full((6, 7), 3.0, order="F"))
    snap = sip.DeviceBlock((6, 7))
    with sip.recording():
        d = m.on_device(sip.WRITE_ON_DEVICE)
        sip.contract_labels([1, 3], [6, 7], [1, 2], dL, [2, 3], dR, out=d)          # recorded, not launched yet
        h = m.on_host(sip.READ_ON_HOST)                                             # must flush, then d2h
        assert np.max(np.abs(h - L @ R)) <= 1e-13
        do = old.on_device(sip.READ_ON_DEVICE)
        snap.scale_and_copy(do, 1.0)                                                # recorded reader of the OLD contents
        old.on_host(sip.WRITE_ON_HOST)[...] = 9.0
        do2 = old.on_device(sip.READ_ON_DEVICE)                                     # h2d of the new contents: after the reader
        assert do2.ptr == do.ptr
        old.destroy()                                                               # free deferred past the recorded ops
    assert np.all(snap.to_numpy() == 3.0)
    m.destroy()


def test_put_initialize_increment_scale(sip):
    """the scalar block ops of SialOpsParallel (sial_ops_parallel.cpp:412-528) at the owner, eager and recorded"""
    A = sip.DistArray([[3, 4], [2, 5]])
    A.track_accesses(True)
    A.put_initialize((2, 2), 1.5)
    A.put_increment((2, 2), 0.25)
    A.put_scale((2, 2), -2.0)
    assert np.all(A.get((2, 2)).to_numpy() == -3.5) and A.get((2, 2)).shape == (4, 5)
    assert np.all(A.get((1, 1)).to_numpy() == 0.0)
    bits = dict(A.section_accesses())
    assert bits[A.block_number((2, 2))] == sip.ACCESS_PUT | sip.ACCESS_PUT_ACCUMULATE | sip.ACCESS_GET
    with sip.recording() as rec:
        for idx in ((1, 1), (1, 2), (2, 1)):
            A.put_initialize(idx, 2.0)
            A.put_increment(idx, 1.0)
    assert rec.stats["launches"] == 2          # three fills in one launch, three increments in the next
    for idx in ((1, 1), (1, 2), (2, 1)):
        assert np.all(A.get(idx).to_numpy() == 3.0)
    A.destroy()
