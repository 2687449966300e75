"""Oracle backend of the SIAL front-end walker (aces4_b200/sial_frontend.py): the same per-block call stream executed
on the host with the CPU oracle (oracle/), blocks as numpy Fortran arrays.  Test infrastructure only."""
import numpy as np


class HostBlock:
    def __init__(self, a):
        self.a = a

    @property
    def shape(self):
        return self.a.shape


class OracleBackend:
    def __init__(self, oracle, arrays, fock=None, moa_seg_ranges=None):
        """arrays: name -> dict {segment tuple: ndarray}; missing blocks of a put target are created as zeros"""
        self.o, self.arrays, self.fock, self.ranges = oracle, arrays, fock, moa_seg_ranges
        self.calls = 0

    def new_block(self, shape):
        return HostBlock(np.full(shape, np.nan, order="F"))

    def free(self, b):
        b.a = None

    def begin_pardo(self):
        pass

    def end_pardo(self):
        pass

    def request(self, name, segs, shape):
        pass

    def array_block(self, name, segs, shape):
        A = self.arrays[name]
        if segs not in A:
            A[segs] = np.zeros(shape, order="F")
        return HostBlock(A[segs])

    def static_block(self, name, segs, shape):
        return self.array_block(name, segs, shape)

    def fill(self, b, v):
        b.a[...] = v
        self.calls += 1

    def scale(self, b, f):
        b.a *= f
        self.calls += 1

    def add_sub(self, d, l, r, sign):
        d.a[...] = l.a + r.a if sign > 0 else l.a - r.a
        self.calls += 1

    def increment(self, b, v):
        b.a += v
        self.calls += 1

    def axpy(self, d, s, f):
        d.a[...] = self.o.block_add(d.a, s.a, f)[0]
        self.calls += 1

    def copy(self, d, dlabs, s, slabs):
        from aces4_b200.sial_frontend import label_numbers
        if tuple(dlabs) == tuple(slabs):
            d.a[...] = s.a
        else:
            dn, sn = label_numbers(dlabs, slabs)
            d.a[...] = self.o.permute_labels(dn, sn, s.a)
        self.calls += 1

    def contract(self, d, dlabs, L, llabs, R, rlabs):
        from aces4_b200.sial_frontend import label_numbers
        dn, ln, rn = label_numbers(dlabs, llabs, rlabs)
        out, ierr = self.o.contract_labels(dn, list(d.shape), ln, L.a, rn, R.a)
        assert ierr == 0
        d.a[...] = out
        self.calls += 1

    def put(self, arr, segs, b):
        self.arrays[arr][segs] = np.array(b.a, order="F")

    def put_accumulate(self, arr, segs, b):
        A = self.arrays[arr]
        if segs not in A:
            A[segs] = np.zeros(b.shape, order="F")
        A[segs] = self.o.block_add(A[segs], b.a, 1.0)[0]

    def put_initialize(self, arr, segs, shape, v):
        self.arrays[arr][segs] = np.full(shape, float(v), order="F")

    def put_increment(self, arr, segs, shape, v):
        A = self.arrays[arr]
        if segs not in A:
            A[segs] = np.zeros(shape, order="F")
        A[segs] = A[segs] + float(v)

    def put_scale(self, arr, segs, shape, f):
        A = self.arrays[arr]
        if segs not in A:
            A[segs] = np.zeros(shape, order="F")
        A[segs] = A[segs] * float(f)

    def has_array(self, name):
        return name in self.arrays

    def host_array(self, b):
        return b.a

    def set_from_host(self, b, a):
        b.a[...] = a

    def block_value(self, b):
        return float(np.asarray(b.a).reshape(-1)[0])

    # persistence: a label registry shared by the backends of consecutive programs (class attribute)
    registry = {}

    def set_persistent(self, name, label):
        self.registry[label] = self.arrays.pop(name)

    def restore_persistent(self, name, label):
        self.arrays[name] = self.registry.pop(label)

    def persist_scalar(self, label, value):
        self.registry[label] = float(value)

    def restore_scalar(self, label):
        return self.registry.pop(label)

    def execute(self, fname, blocks, segs, kinds, bare):
        if fname == "stripi":
            y, ierr = self.o.si_stripi(np.asfortranarray(blocks[0].a), list(segs[0]), blocks[1].a.shape, list(segs[1]), self.ranges)
            assert ierr == 0, ierr
            blocks[1].a[...] = y
            self.calls += 1
            return
        if fname == "energy_ty_denominator_rhf":
            assert self.o.si_energy_ty_denominator_rhf(blocks[0].a, list(segs[0]), self.fock, float(bare[1]), self.ranges) == 0
            self.calls += 1
            return
        if fname in ("anti_symm_o", "anti_symm_v"):
            assert getattr(self.o, "si_" + fname)(blocks[0].a, list(segs[0]), self.ranges) == 0
            self.calls += 1
            return
        if fname == "invert_diagonal":
            assert self.o.si_invert_diagonal(blocks[0].a, blocks[1].a) == 0
            self.calls += 1
            return
        if fname == "invert_diagonal_asym":
            assert self.o.si_invert_diagonal_asym(blocks[0].a, list(segs[0]), blocks[1].a, self.ranges) == 0
            self.calls += 1
            return
        if fname == "return_diagonal_elements":
            assert self.o.si_return_diagonal_elements(blocks[0].a, list(segs[0]), self.ranges) == 0
            self.calls += 1
            return
        assert fname == "energy_denominator_rhf", fname
        assert self.o.si_energy_denominator_rhf(blocks[0].a, list(segs[0]), self.fock, self.ranges) == 0

    def reshaped(self, b, shape):
        return b if tuple(b.a.shape) == tuple(shape) else HostBlock(b.a.reshape(shape, order="F"))

    def moa_seg_ranges(self):
        return list(self.ranges)

    def dot(self, L, llabs, R, rlabs, prev):
        from aces4_b200.sial_frontend import label_numbers
        a = L.a
        if tuple(llabs) != tuple(rlabs):
            ln, rn = label_numbers(llabs, rlabs)
            a = self.o.permute_labels(rn, ln, a)
        return float(np.sum(a * R.a))

    def scalar_set(self, prev, v):
        return float(v)

    def scalar_axpy(self, s, f, other):
        return s + f if other is None else s + f * other

    def scalar_scale(self, s, f):
        return s * f

    def barrier(self):
        pass

    def collective_sum(self, a, b):
        return a + b

    def value(self, s):
        return float(s)
