"""A numpy stand-in for the PYTHON surface of aces4_b200.api that the SIAL front-end's DeviceBackend and the device
energy tests touch -- blocks as Fortran-ordered numpy arrays, the arithmetic done by the CPU oracle.  TEST
INFRASTRUCTURE: it exists so that the host-side logic of tests that have not yet seen a GPU (DeviceBackend's scalar
handling across iterations, local arrays, the test bodies themselves) can be executed here, numerically, on the CPU
(tests/test_device_tests_on_fake_api_cpu.py).  It says nothing about libsipgpu itself; nothing in aces4_b200/ imports it."""
import itertools

import numpy as np


class _Lib:
    def __init__(self, api):
        self.api = api

    def sipgpu_h2d(self, ptr, host, n):
        blk = self.api._by_ptr[ptr]
        blk.a[...] = np.asarray(host).reshape(blk.a.shape, order="F")
        return 0

    def sipgpu_block_dot_accumulate(self, lptr, rptr, n, accptr):
        L, R, acc = (self.api._by_ptr[p] for p in (lptr, rptr, accptr))
        acc.a[...] += float(np.sum(L.a * R.a))
        self.api._launches += 1
        return 0


class FakeApi:
    def __init__(self, oracle):
        self.o = oracle
        self._by_ptr = {}
        self._ptr = itertools.count(0x1000, 0x1000)
        self._launches = 0
        self._moa = None
        self._registry = {}
        self._lib = _Lib(self)
        api = self

        class DeviceBlock:
            def __init__(self, shape, zero=False, a=None, ptr=None, owned=True):
                self.shape = tuple(int(x) for x in shape)
                if ptr is not None:      # a view of an existing block under other extents (same elements)
                    a = api._by_ptr[ptr].a.reshape(self.shape, order="F")
                self.a = a if a is not None else np.full(self.shape, 0.0 if zero else np.nan, order="F")
                self.size, self.rank = int(self.a.size), len(self.shape)
                self.ptr = next(api._ptr)
                api._by_ptr[self.ptr] = self

            @classmethod
            def from_numpy(cls, a):
                return cls(a.shape, a=np.array(a, order="F", dtype=float))

            def to_numpy(self):
                return np.array(self.a, order="F")

            def free(self):
                api._by_ptr.pop(self.ptr, None)
                self.a = None

            def _op(self):
                api._launches += 1
                return self

            def fill(self, v):
                self.a[...] = v
                return self._op()

            def scale(self, f):
                self.a *= f
                return self._op()

            def increment(self, d):
                self.a += d
                return self._op()

            def axpy(self, other, f):
                self.a[...] = api.o.block_add(self.a, other.a, f)[0]
                return self._op()

            def set_add_sub(self, l, r, sign):
                self.a[...] = l.a + r.a if sign > 0 else l.a - r.a
                return self._op()

            def accumulate(self, other):
                return self.axpy(other, 1.0)

            def scale_and_copy(self, other, f):
                self.a[...] = f * other.a
                return self._op()

        class DistArray:
            def __init__(self, seg_ext_per_index, my_rank=0, world=1, exchange=None, devices=None):
                assert world == 1
                self.seg_ext = [list(map(int, s)) for s in seg_ext_per_index]
                self.blocks = {}

            def block_shape(self, idx):
                return tuple(self.seg_ext[d][i - 1] for d, i in enumerate(idx))

            def owner(self, idx):
                return 0

            def block_view(self, idx):
                idx = tuple(int(i) for i in idx)
                if idx not in self.blocks:
                    self.blocks[idx] = DeviceBlock(self.block_shape(idx), zero=True)
                return self.blocks[idx]

            def get(self, idx, out=None):
                return DeviceBlock.from_numpy(self.block_view(idx).a)

            def put(self, idx, blk):
                self.block_view(idx).scale_and_copy(blk, 1.0)

            def put_accumulate(self, idx, blk):
                self.block_view(idx).accumulate(blk)

            def put_increment(self, idx, d):
                self.block_view(idx).increment(d)

            def put_scale(self, idx, f):
                self.block_view(idx).scale(f)

            def put_initialize(self, idx, v):
                self.block_view(idx).fill(v)

            def fill_local(self, v):
                for b in self.blocks.values():
                    b.fill(v)

            def destroy(self):
                self.blocks = None

            def persist(self, label):
                api._registry[label] = (self.seg_ext, self.blocks)
                self.blocks = None

            def restore(self, label):
                seg_ext, blocks = api._registry.pop(label)
                assert seg_ext == self.seg_ext, "restore_persistent into an array of another layout"
                self.blocks = blocks

        self.DeviceBlock, self.DistArray = DeviceBlock, DistArray

    # ---- module-level functions of aces4_b200.api ----
    def lib(self):
        return self._lib

    @staticmethod
    def _check(rc, what=""):
        assert rc == 0, what

    @staticmethod
    def _hp(a):
        return a

    def sync(self):
        pass

    def kernel_launches(self):
        return self._launches

    def wl_begin(self, dry=False):
        pass

    def wl_end(self):
        return {}

    def persist_scalar(self, label, value):
        self._registry[label] = float(value)

    def restore_scalar(self, label):
        return self._registry.pop(label)

    def set_predefined_int_array(self, name, values):
        assert name == "moa_seg_ranges"
        self._moa = list(values)

    def permute_labels(self, lhs_labels, rhs_labels, rhs, out=None):
        res = self.o.permute_labels(list(lhs_labels), list(rhs_labels), rhs.a)
        if out is None:
            return self.DeviceBlock.from_numpy(res)
        out.a[...] = res
        self._launches += 1
        return out

    def contract_labels(self, dlab, dext, llab, L, rlab, R, out=None, alpha=1.0, beta=0.0):
        res, ierr = self.o.contract_labels(list(dlab), list(dext), list(llab), L.a, list(rlab), R.a)
        assert ierr == 0
        if out is None:
            out = self.DeviceBlock(tuple(dext), zero=True)
        out.a[...] = alpha * res.reshape(out.a.shape, order="F") + (beta * out.a if beta != 0.0 else 0.0)
        self._launches += 1
        return out

    def get_contraction_ptrn(self, dlab, llab, rlab):
        return self.o.get_contraction_ptrn(list(dlab), list(llab), list(rlab))

    def contract_sliced(self, ptrn, L, lext, lbeg, R, rext, rbeg, dext, out=None, dbeg=None, alpha=1.0, beta=0.0):
        assert dbeg is None and alpha == 1.0 and beta == 0.0

        def operand(blk, ext, beg):
            if beg is None:
                return blk.a
            return np.asfortranarray(blk.a[tuple(slice(b, b + e) for b, e in zip(beg, ext))])

        res, ierr = self.o.block_contract(list(ptrn), operand(L, lext, lbeg), operand(R, rext, rbeg), tuple(dext))
        assert ierr == 0
        out.a[...] = res.reshape(out.a.shape, order="F")
        self._launches += 1
        return out

    def si_stripi(self, x, iv0, y, iv1):
        self._launches += 1
        out, ierr = self.o.si_stripi(np.asfortranarray(x.a), list(iv0), y.a.shape, list(iv1), self._moa)
        if ierr == 0:
            y.a[...] = out
        return ierr

    def si_anti_symm_o(self, block, index_values):
        self._launches += 1
        return self.o.si_anti_symm_o(block.a, list(index_values), self._moa)

    def si_anti_symm_v(self, block, index_values):
        self._launches += 1
        return self.o.si_anti_symm_v(block.a, list(index_values), self._moa)

    def si_return_diagonal_elements(self, block, index_values):
        self._launches += 1
        return self.o.si_return_diagonal_elements(block.a, list(index_values), self._moa)

    def si_invert_diagonal(self, a1, a2):
        self._launches += 1
        return self.o.si_invert_diagonal(a1.a, a2.a)

    def si_invert_diagonal_asym(self, a1, index_values, a2):
        self._launches += 1
        return self.o.si_invert_diagonal_asym(a1.a, list(index_values), a2.a, self._moa)

    def si_energy_ty_denominator_rhf(self, block, index_values, fock, shift_block):
        self._launches += 1
        return self.o.si_energy_ty_denominator_rhf(block.a, list(index_values), fock.a, float(shift_block.a.reshape(-1)[0]), self._moa)

    def si_energy_denominator_rhf(self, block, index_values, fock):
        self._launches += 1
        return self.o.si_energy_denominator_rhf(block.a, list(index_values), fock.a, self._moa)
