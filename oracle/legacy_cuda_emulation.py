"""numpy restatement of the reference's legacy CUDA contraction / permutation helpers, statement by statement:
`__gpu_contract_helper` and `__gpu_permute_helper`, src/sip/cuda/gpu_super_instructions.cu:369-586 / 588-684, with the
`reorderScatter` / `reorderGather` kernels (:307-353).  TEST INFRASTRUCTURE ONLY.

Why it exists: tests/test_gpu_vs_ref_cuda.py compares the product with the real thing (that file compiled for sm_100a) on a
GPU; this emulation lets the CPU suite check beforehand that the legacy ALGORITHM, as written, agrees with the oracle on
every case that test will run -- so a disagreement on the GPU can only come from the product or from the legacy code's
execution, not from the choice of cases (e.g. a label order the legacy index bookkeeping does not handle).
"""
import numpy as np

MAX_RANK = 6


def _walk(dims, steps, n, size):
    """the index arithmetic shared by both kernels (:315-325 / :339-349): linear index -> sum of digit_i * steps[i]"""
    t = np.arange(size, dtype=np.int64)
    out = np.zeros(size, dtype=np.int64)
    for i in range(n):
        out += (t % dims[i]) * steps[i]
        t //= dims[i]
    return out


def reorder_scatter(old, dims, steps, n, size):        # newX[newIndex] = oldX[oldIndex]
    new = np.zeros(size)
    new[_walk(dims, steps, n, size)] = old[:size]
    return new


def reorder_gather(old, dims, steps, n, size):         # newX[newIndex] = oldX[oldIndex(newIndex)]
    return old[_walk(dims, steps, n, size)]


def gpu_contract(y_dims, y_inds, x1, x1_dims, x1_inds, x2, x2_dims, x2_inds):
    ny, n1, n2 = len(y_dims), len(x1_dims), len(x2_dims)
    nc = (n1 + n2 - ny) // 2                                                     # :422
    x1IP, x1DP, x2IP, x2DP, yIP, yDP = ([None] * MAX_RANK for _ in range(6))
    c = k = 0
    for i in range(n1):                                                          # :426-450
        contracted = False
        for j in range(n2):
            if x1_inds[i] == x2_inds[j]:
                contracted = True
                x1IP[n1 - nc + c], x1DP[n1 - nc + c] = x1_inds[i], x1_dims[i]
                x2IP[c], x2DP[c] = x2_inds[j], x2_dims[j]
                c += 1
                break
        if not contracted:
            x1IP[k], x1DP[k], yIP[k], yDP[k] = x1_inds[i], x1_dims[i], x1_inds[i], x1_dims[i]
            k += 1
    c = 0
    for i in range(n2):                                                          # :452-464
        for j in range(ny):
            if x2_inds[i] == y_inds[j]:
                x2IP[nc + c], x2DP[nc + c] = x2_inds[i], x2_dims[i]
                yIP[k], yDP[k] = y_inds[j], y_dims[j]
                k += 1
                c += 1

    def steps_for(inds, indsP, dimsP, n):                                        # :467-476, :496-505, :553-562
        steps, step = [0] * MAX_RANK, 1
        for i in range(n):
            for j in range(n):
                if inds[j] == indsP[i]:
                    steps[j] = step
                    break
            step *= dimsP[i]
        return steps, step

    steps, size = steps_for(x1_inds, x1IP, x1DP, n1)
    s1 = reorder_scatter(np.ravel(x1, order="F"), x1_dims, steps, n1, size)
    steps, size = steps_for(x2_inds, x2IP, x2DP, n2)
    s2 = reorder_scatter(np.ravel(x2, order="F"), x2_dims, steps, n2, size)
    lda = int(np.prod([x1DP[i] for i in range(n1 - nc)], dtype=np.int64))        # :526-532
    ldb = int(np.prod([x2DP[i] for i in range(nc)], dtype=np.int64))
    a = s1.reshape((lda, ldb), order="F")                                        # cublasDgemm(N, N, lda, size/ldb, ldb) :536
    b = s2.reshape((ldb, size // ldb), order="F")
    s3 = np.ravel(a @ b, order="F")
    steps, size = steps_for(y_inds, yIP, yDP, ny)
    return reorder_gather(s3, y_dims, steps, ny, size).reshape(tuple(y_dims), order="F")   # :567-573


def gpu_permute(y_dims, y_inds, x1, x1_dims, x1_inds):
    ny, n1 = len(y_dims), len(x1_dims)
    steps, step = [0] * MAX_RANK, 1                                              # :658-667 (x1IndsP = x1Inds, x1DimsP = x1Dims, :631-637)
    for i in range(ny):
        for j in range(ny):
            if y_inds[j] == x1_inds[i]:
                steps[j] = step
                break
        step *= x1_dims[i]
    return reorder_gather(np.ravel(x1, order="F"), y_dims, steps, ny, step).reshape(tuple(y_dims), order="F")
