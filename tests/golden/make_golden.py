#!/usr/bin/env python
"""Regenerate tests/golden/fixtures.npz from the reference's literal test matrices.

Runs ONLY in the build container (needs /root/reference, which does not exist on the GPU box).
Source: /root/reference/test/matrices.jl:4-9 -- six SuiteSparse matrices written as Julia
`sparse(I, J, V, m, n)` literals (the sixth wrapped in `Symmetric(..., Symbol("L"))`), which
test/runtests.jl:12-18 feeds through `SparseMatrixCSC(A)` before building the VBC types.

What is stored per fixture (keys `<name>/<field>`):
  m, n            dimensions
  colptr, rowval  canonical CSC, 1-based int64, rows ascending inside a column (what
                  Julia's `sparse` + `SparseMatrixCSC(Symmetric(...))` produce)
  nzval           float64 values (Bool fixtures become 0.0/1.0; `sparse` sums duplicates --
                  for Bool that is `|`, i.e. still 1)
  is_bool         1 if the Julia literal was `Bool[...]`

Usage:  python tests/golden/make_golden.py
"""
import os
import re
import sys

import numpy as np

REF = "/root/reference/test/matrices.jl"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures.npz")


def _parse_vec(s):
    is_bool = s.startswith("Bool[")
    body = s[s.index("[") + 1 : s.rindex("]")]
    arr = np.array([float(t) for t in body.split(",")], dtype=np.float64)
    return arr, is_bool


def _split_top_level(s):
    """Split 'a, [b, c], d' on commas that are outside brackets."""
    out, depth, cur = [], 0, []
    for ch in s:
        if ch in "[(":
            depth += 1
        elif ch in "])":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    if cur:
        out.append("".join(cur).strip())
    return out


def to_csc(I, J, V, m, n, is_bool):
    """Julia `sparse(I,J,V,m,n)`: column-major, rows sorted, duplicates combined (+ or |)."""
    I = I.astype(np.int64)
    J = J.astype(np.int64)
    order = np.lexsort((I, J))
    I, J, V = I[order], J[order], V[order]
    keep = np.ones(len(I), dtype=bool)
    keep[1:] = (I[1:] != I[:-1]) | (J[1:] != J[:-1])
    grp = np.cumsum(keep) - 1
    Vc = np.zeros(int(keep.sum()))
    np.add.at(Vc, grp, V)
    if is_bool:
        Vc = (Vc != 0).astype(np.float64)
    I, J = I[keep], J[keep]
    colptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(J - 1, minlength=n))]).astype(np.int64)
    return colptr, I.copy(), Vc


def main():
    if not os.path.exists(REF):
        sys.exit("reference not mounted; fixtures.npz is committed, nothing to do")
    out = {}
    names = []
    for line in open(REF).read().splitlines():
        mt = re.match(r'^"([^"]+)" => (.*),\s*$', line)
        if not mt:
            continue
        name, expr = mt.group(1), mt.group(2)
        sym = expr.startswith("Symmetric(")
        if sym:
            assert expr.endswith('Symbol("L"))')
            expr = expr[len("Symmetric(") : expr.rindex(", Symbol")]
        assert expr.startswith("sparse(") and expr.endswith(")")
        args = _split_top_level(expr[len("sparse(") : -1])
        assert len(args) == 5, len(args)
        I, _ = _parse_vec(args[0])
        J, _ = _parse_vec(args[1])
        V, is_bool = _parse_vec(args[2])
        m, n = int(args[3]), int(args[4])
        if sym:
            # Symmetric(A, :L): the lower triangle defines the matrix (runtests.jl:18
            # materialises it with SparseMatrixCSC(A)).
            low = I >= J
            I, J, V = I[low], J[low], V[low]
            off = I != J
            I, J, V = (np.concatenate([I, J[off]]), np.concatenate([J, I[off]]),
                       np.concatenate([V, V[off]]))
        colptr, rowval, nzval = to_csc(I, J, V, m, n, is_bool)
        key = name.replace("/", "__")
        names.append(key)
        out[f"{key}/m"] = np.int64(m)
        out[f"{key}/n"] = np.int64(n)
        out[f"{key}/colptr"] = colptr
        out[f"{key}/rowval"] = rowval
        out[f"{key}/nzval"] = nzval
        out[f"{key}/is_bool"] = np.int64(is_bool)
        print(f"{name}: {m}x{n} nnz={len(rowval)} bool={is_bool} sym={sym}")
    out["names"] = np.array(names)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
