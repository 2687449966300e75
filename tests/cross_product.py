"""Config 4 of BASELINE.json / SURVEY.md section 8(d) item 2: the FULL cross product of rank-4 contractions
D[p0,p1,p2,p3] = L[..]*R[..] with two contracted indices -- 24 destination permutations x 6 placements of the two
contracted labels in L x 6 in R x 2 relative orders of the contracted pair = 1728 label patterns."""
import itertools

FREE_L, FREE_R, CONTRACTED = (1, 2), (3, 4), (5, 6)


def _placements(free, pair):
    for pos in itertools.combinations(range(4), 2):
        labs, f, c = [0] * 4, iter(free), iter(pair)
        for i in range(4):
            labs[i] = next(c) if i in pos else next(f)
        yield labs


def patterns():
    """yields (dlab, llab, rlab), 1728 of them, all distinct"""
    for llab in _placements(FREE_L, CONTRACTED):
        for pair in (CONTRACTED, CONTRACTED[::-1]):
            for rlab in _placements(FREE_R, pair):
                for dlab in itertools.permutations(FREE_L + FREE_R):
                    yield list(dlab), list(llab), list(rlab)


def einsum_spec(dlab, llab, rlab):
    s = "abcdefgh"
    return "".join(s[x] for x in llab) + "," + "".join(s[x] for x in rlab) + "->" + "".join(s[x] for x in dlab)
