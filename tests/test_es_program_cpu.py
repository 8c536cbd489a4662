"""CPU-only check of the Elston-Stewart pedigree compiler (famseq_b200/csrc/host/es_program.cpp).

The compiled message program is fetched through the C ABI (fs_get_es_program, host-only engine) and interpreted
here, in plain Python floats (IEEE doubles, no FMA), with the semantics documented in es_program.hpp -- the same
semantics cuda/es_kernel.cu implements.  The result must be bit-identical to the oracle's recursive evaluation.
This is test infrastructure: the product has no CPU compute path."""
import numpy as np
import pytest

import famseq_b200 as fs
from famseq_b200 import synth
from oracle import oracle as O

OP_END, OP_MUL, OP_ANT, OP_POS, OP_FIN = 0, 1, 2, 3, 4
TAB_AUTO, TAB_XF, TAB_XM = 0, 1, 2


def interpret(words, n_slots, tabs, priors, lk, flag):
    """One variant.  lk: [S][3].  Returns (post [S][3] or None when a row sum is zero)."""
    known, chrx = flag & 1, (flag >> 1) & 1
    pa = priors[1 if known else 0]
    pm = priors[(3 if known else 2)] if chrx else pa
    slots = [[0.0] * 3 for _ in range(n_slots)]
    post = [[0.0] * 3 for _ in range(len(lk))]

    S = len(lk)
    # vector file: [0,S) likelihood rows, [S,S+n_slots) scratch, then prior (non-male), prior (male), ones
    slots = [[float(x) for x in row] for row in lk] + slots + [[float(x) for x in pa], [float(x) for x in pm], [1.0, 1.0, 1.0]]

    def load(u):
        return list(slots[u])

    def T(male, g, a, b):
        t = tabs[(TAB_XM if male else TAB_XF) if chrx else TAB_AUTO]
        return float(t[g * 9 + a * 3 + b])

    failed, pc = False, 0
    while True:
        w0 = int(words[pc]); op = w0 & 0xff
        if op == OP_END:
            break
        if op == OP_MUL:
            w1 = int(words[pc + 1]); a, b = load(w1 & 0xffff), load(w1 >> 16)
            slots[(w0 >> 8) & 0xffff] = [a[g] * b[g] for g in range(3)]
            pc += 2
        elif op == OP_ANT:
            w1 = int(words[pc + 1]); nsib = w0 >> 25; cmale = (w0 >> 24) & 1
            wm, wf = load(w1 & 0xffff), load(w1 >> 16)
            sibs = None
            for k in range(nsib):
                wk = int(words[pc + 2 + k]); d = load(wk & 0xffff); smale = (wk >> 16) & 1
                cur = [[(d[0] * T(smale, 0, a, b) + d[1] * T(smale, 1, a, b)) + d[2] * T(smale, 2, a, b) for b in range(3)] for a in range(3)]
                sibs = cur if sibs is None else [[sibs[a][b] * cur[a][b] for b in range(3)] for a in range(3)]
            out = []
            for g in range(3):
                over_m = 0.0
                for a in range(3):
                    over_f = 0.0
                    for b in range(3):
                        term = wf[b] * T(cmale, g, a, b)
                        if nsib:
                            term = term * sibs[a][b]
                        over_f = over_f + term
                    over_m = over_m + wm[a] * over_f
                out.append(over_m)
            slots[(w0 >> 8) & 0xffff] = out
            pc += 2 + nsib
        elif op == OP_POS:
            w1 = int(words[pc + 1]); nkid = w0 >> 25; i_second = bool(chrx and ((w0 >> 24) & 1))
            wj = load(w1 & 0xffff)
            kids = None
            for k in range(nkid):
                wk = int(words[pc + 2 + 2 * k]); kmale = int(words[pc + 3 + 2 * k]) & 1
                lkc, mc = load(wk & 0xffff), load(wk >> 16)
                cur = [[0.0] * 3 for _ in range(3)]
                for g in range(3):
                    for b in range(3):
                        sc = None
                        for l in range(3):
                            tr = T(kmale, l, b, g) if i_second else T(kmale, l, g, b)
                            term = (tr * lkc[l]) * mc[l]
                            sc = term if sc is None else sc + term
                        cur[g][b] = sc
                kids = cur if kids is None else [[kids[g][b] * cur[g][b] for b in range(3)] for g in range(3)]
            slots[(w0 >> 8) & 0xffff] = [(wj[0] * kids[g][0] + wj[1] * kids[g][1]) + wj[2] * kids[g][2] for g in range(3)]
            pc += 2 + 2 * nkid
        else:
            w1, w2 = int(words[pc + 1]), int(words[pc + 2])
            m, l, a = load(w1 & 0xffff), load(w1 >> 16), load(w2 & 0xffff)
            v = [(m[g] * l[g]) * a[g] for g in range(3)]
            s = (v[0] + v[1]) + v[2]
            if s == 0.0:
                failed = True
            elif (w0 >> 8) & 1:
                post[w0 >> 9] = [v[g] / s for g in range(3)]
            pc += 3
    return None if failed else post


@pytest.mark.parametrize("name", ["trio", "ped14", "half_sibs", "three_wives"])
def test_compiled_program_reproduces_the_oracle_bit_for_bit(name):
    ped = synth.PEDIGREES[name]()
    cols = ped.sequenced_cols()
    V = 40
    lk, fl = synth.synth_likelihoods(ped, V, seed=123, x_fraction=0.4)
    want = O.run(ped, cols, lk, fl, method=O.ES, lc=0.0 + 5.0)  # -LRC 5: the pedigree is always used
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, device=-1) as e:
        words, n_slots = e.es_program()
        a, xf, xm, _, _ = e.tables()
        priors = e.params.priors()
    for v in range(V):
        got = interpret(words, n_slots, [a, xf, xm], priors, lk[v], int(fl[v]))
        if want["status"][v]:
            continue
        assert got is not None
        assert np.array_equal(np.array(got), want["post"][v]), f"{name} variant {v}"


def test_partial_sequencing_program():
    ped = synth.ped14()
    cols = [13, 2, 7, 0, 10, 5]
    lk, fl = synth.synth_likelihoods(synth._mk([(i, 0, 0, 1) for i in range(1, 7)]), 30, seed=9, x_fraction=0.3)
    want = O.run(ped, cols, lk, fl, method=O.ES, lc=5.0)
    with fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, cols, device=-1) as e:
        words, n_slots = e.es_program()
        a, xf, xm, _, _ = e.tables()
        priors = e.params.priors()
    for v in range(30):
        got = interpret(words, n_slots, [a, xf, xm], priors, lk[v], int(fl[v]))
        if not want["status"][v]:
            assert np.array_equal(np.array(got), want["post"][v])
