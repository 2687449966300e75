"""CPU restatement of the synthetic CCSD iteration (aces4_b200/sial_workload.py) for parity tests: the same seeded
blocks (oracle.fill_hash), assembled into dense tensors, contracted with numpy.einsum, and per destination block with
the oracle's tensor_block_contract chain (contract into a temp, then +=, like the SIAL body).  Test infrastructure."""
import numpy as np

from aces4_b200.sial_workload import SCALE, TAG, TERMS, block_present


def _offsets(segs):
    o = [0]
    for s in segs:
        o.append(o[-1] + s)
    return o


class RefWorkload:
    def __init__(self, oracle, o_segs, v_segs, seed=0xACE54, ao_pool=8, density=1.0):
        self.oracle, self.o_segs, self.v_segs, self.seed, self.ao_pool = oracle, list(o_segs), list(v_segs), seed, ao_pool
        self.density = density
        self.segs = {"v": self.v_segs, "o": self.o_segs}
        kinds = {"T2old": "vovo", "Vvovo": "vovo", "Voooo": "oooo", "TY": "vovo", "Vovvo": "ovvo", "Vvvoo": "vvoo"}
        self.kinds = kinds
        self.blocks = {}
        self.dense = {}
        for name, kind in kinds.items():
            self.dense[name] = self._assemble(name, kind)
        T = self.dense["T2old"]
        self.dense["W"] = T - np.transpose(T, (0, 3, 2, 1))  # W[c,k,a,i] = T[c,k,a,i] - T[c,i,a,k]
        self.dense["aoint"] = self._assemble_ao()

    def _tag(self, name, number):
        return (TAG[name] << 40) | int(number)

    def block(self, name, idx):
        kind = self.kinds[name]
        nseg = [len(self.segs[k]) for k in kind]
        number = 0
        for p in range(len(idx)):
            number = number * nseg[p] + (idx[p] - 1)  # last index fastest (array_table.cpp:50-97)
        shape = tuple(self.segs[k][i - 1] for k, i in zip(kind, idx))
        if name == "T2old" and not block_present(self.seed, number, getattr(self, "density", 1.0)):
            return np.zeros(shape, order="F")   # an absent block of the block-sparse amplitude array
        return self.oracle.fill_hash(shape, self.seed, self._tag(name, number), SCALE[name])

    def ao_block(self, lam, mu, sig, nu):
        ext = (self.v_segs[lam - 1], self.v_segs[mu - 1], self.v_segs[sig - 1], self.v_segs[nu - 1])
        slot = (lam * 7 + mu * 3 + sig * 5 + nu) % self.ao_pool
        return self.oracle.fill_hash(ext, self.seed, self._tag("aoint", slot * 1000003 + hash(ext) % 1000003), SCALE["aoint"])

    def _assemble(self, name, kind):
        offs = [_offsets(self.segs[k]) for k in kind]
        full = np.zeros([o[-1] for o in offs], order="F")
        for idx in np.ndindex(*[len(self.segs[k]) for k in kind]):
            idx1 = tuple(i + 1 for i in idx)
            sl = tuple(slice(offs[d][idx[d]], offs[d][idx[d] + 1]) for d in range(len(kind)))
            full[sl] = self.block(name, idx1)
        return full

    def _assemble_ao(self):
        offs = _offsets(self.v_segs)
        n = len(self.v_segs)
        full = np.zeros([offs[-1]] * 4, order="F")
        for idx in np.ndindex(n, n, n, n):
            sl = tuple(slice(offs[idx[d]], offs[idx[d] + 1]) for d in range(4))
            full[sl] = self.ao_block(*(i + 1 for i in idx))
        return full

    def iterate(self, terms=None):
        """Dense T2new and the energy scalar."""
        v, o = sum(self.v_segs), sum(self.o_segs)
        direct = np.zeros((v, o, v, o), order="F")
        xs = 0.5 * self.dense["Vvovo"]
        for t in TERMS:
            if terms is not None and t["name"] not in terms:
                continue
            r = t["alpha"] * np.einsum(f"{t['llab']},{t['rlab']}->{t['dlab']}", self.dense[t["L"]], self.dense[t["R"]],
                                       optimize=True)
            if t["sym"]:
                xs = xs + r
            else:
                direct = direct + r
        t2new = direct + xs + np.transpose(xs, (2, 3, 0, 1))
        V = self.dense["Vvovo"]
        energy = float(np.sum(t2new * (2.0 * V - np.transpose(V, (0, 3, 2, 1)))))
        return t2new, energy

    def dest_block_by_oracle(self, term, blk):
        """One destination block of one term through the oracle's block contraction, chained over the contracted
        segments exactly as the SIAL body does (contract into a temp, accumulate)."""
        dlab, llab, rlab = term["dlab"], term["llab"], term["rlab"]
        labs = sorted(set(dlab + llab + rlab))
        num = {c: n + 1 for n, c in enumerate(labs)}
        isv = lambda c: c in "abcd"
        contracted = [c for c in llab if c in rlab]
        segs = dict(zip(dlab, blk))
        dext = [(self.v_segs if isv(c) else self.o_segs)[segs[c] - 1] for c in dlab]
        acc = np.zeros(dext, order="F")
        for cseg in np.ndindex(*[len(self.v_segs if isv(c) else self.o_segs) for c in contracted]):
            for c, s in zip(contracted, cseg):
                segs[c] = s + 1
            ops = []
            for name, lab in ((term["L"], llab), (term["R"], rlab)):
                idx = tuple(segs[c] for c in lab)
                if name == "aoint":
                    ops.append(self.ao_block(*idx))
                elif name == "W":
                    c_, k_, a_, i_ = idx
                    ops.append(self.block("T2old", idx) - np.transpose(self.block("T2old", (c_, i_, a_, k_)), (0, 3, 2, 1)))
                else:
                    ops.append(self.block(name, idx))
            d, ierr = self.oracle.contract_labels([num[c] for c in dlab], dext, [num[c] for c in llab], ops[0],
                                                  [num[c] for c in rlab], ops[1])
            assert ierr == 0
            acc += d
        return term["alpha"] * acc
