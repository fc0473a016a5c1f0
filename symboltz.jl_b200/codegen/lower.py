"""Lowering of the symbolic model to CUDA C (the code generator proper).

What the reference does at `CosmologyProblem` time (src/solve.jl:129-236: mtkcompile, analytic sparse
Jacobian `jac=true, sparse=true`, background->spline substitution) is done here with sympy, B200-first:

 * The perturbation system is linear in u, f(u,τ) = J(τ)·u, so only J is lowered.  J is split as
       J(τ) = J_local(τ) + p(τ)·φ(τ)ᵀ + q(τ)·ψ(τ)ᵀ
   where Φ̇ = φᵀu and Ψ = ψᵀu are the two "hub" observed variables that make the Jacobian's dense
   arrowhead rows/columns.  J_local is what remains: tridiagonal hierarchy chains plus a few local
   couplings.  Every nonzero is `constant × basis_m(τ,k)` with a small set of basis functions b_m that
   the generated device function `sb_basis` evaluates from the splined background.
 * A symbolic factorisation of the fixed pattern of J_local (leaf-peeling elimination forest: zero
   fill; what is left are small dense 2-core blocks, pivoted at run time) is emitted as tables; the
   integrator kernel factors W = I/(γh) − J with it and handles the two hubs with a rank-2 Woodbury
   correction.  This replaces KLU (src/solve.jl:329).
 * ∂f/∂τ = J'(τ)u uses the same tables with ḃ_m (chain rule through the spline derivative).
 * Background RHS g(y) and its analytic 5x5 Jacobian (host solver), κ̈ and κ⃛ by flow differentiation
   (for the CMB sources), closed-form initial conditions and Δm are emitted as straight-line code.
"""
import hashlib
import sympy as sp
from sympy.printing.c import C99CodePrinter

from .model import Model


class _Printer(C99CodePrinter):
    def _print_Integer(self, expr):  # always floating literals (sqrt(2) would otherwise bind to std::sqrt<int> on the host)
        return "%d.0" % int(expr)

    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer and 2 <= abs(int(e)) <= 4:
            s = "(" + "*".join(["(%s)" % self._print(b)] * abs(int(e))) + ")"
            return s if e > 0 else "(1.0/%s)" % s
        if e == sp.Rational(1, 2):
            return "sqrt(%s)" % self._print(b)
        if e == -sp.Rational(1, 2):
            return "(1.0/sqrt(%s))" % self._print(b)
        if e == -1:
            return "(1.0/(%s))" % self._print(b)
        return super()._print_Pow(expr)


_pr = _Printer()


def _ccode(e):
    return _pr.doprint(e)


def _needed(stmts, targets):
    """Backward reachability: statements required to evaluate the symbols in `targets`."""
    need = set()
    for t in targets:
        need |= sp.sympify(t).free_symbols
    keep = []
    for s, e in reversed(stmts):
        if s in need:
            keep.append((s, e))
            need |= e.free_symbols
    return list(reversed(keep))


def tangent(stmts, seeds, prefix):
    """Forward-mode differentiation of a straight-line program.
    seeds: {symbol: direction expression} for inputs (and possibly temporaries). Statements whose
    tangent vanishes identically are dropped. Returns (tangent statements, {symbol: tangent symbol or 0})."""
    d = dict(seeds)
    out = []
    for s, e in stmts:
        acc = sp.Integer(0)
        for a in e.free_symbols:
            da = d.get(a, 0)
            if da != 0:
                acc += sp.diff(e, a) * da
        if acc == 0:
            d[s] = sp.Integer(0)
        else:
            ds = sp.Symbol(prefix + str(s), real=True)
            out.append((ds, acc))
            d[s] = ds
    return out, d


def _merge(*stmt_lists):
    """Concatenate programs that may share statements (same symbol, same expr), keeping first occurrences."""
    seen, out = set(), []
    for lst in stmt_lists:
        for s, e in lst:
            if s not in seen:
                seen.add(s)
                out.append((s, e))
    return out


def _emit(stmts, outputs, subs, indent="    "):
    """stmts: program; outputs: list of (lhs string, expr). Prunes, prints. Returns (code, op count)."""
    stmts = _needed(stmts, [e for _, e in outputs])
    lines, nops = [], 0
    for s, e in stmts:
        lines.append(f"{indent}const double {s} = {_ccode(e.xreplace(subs))};")
        nops += int(sp.count_ops(e))
    for lhs, e in outputs:
        e = sp.sympify(e)
        lines.append(f"{indent}{lhs} = {_ccode(e.xreplace(subs))};")
        nops += int(sp.count_ops(e))
    return "\n".join(lines), nops


class Lowered:
    pass


def _linear_terms(expr, syms):
    """expr is an expanded sum of terms each linear in exactly one symbol of `syms`. Returns {sym: coeff}."""
    out = {}
    symset = set(syms)
    for term in sp.Add.make_args(expr):
        hit = term.free_symbols & symset
        assert len(hit) == 1, f"term {term} not linear in exactly one unknown"
        s = hit.pop()
        out[s] = out.get(s, 0) + term / s
    return out


def analyse(m: Model):
    """Hub split, basis extraction and symbolic factorisation. Pure Python/sympy; no code emitted yet."""
    L = Lowered()
    N = m.N
    k = m.k
    Phd, Psi = m.hubs
    uidx = {s: i for i, s in enumerate(m.u)}
    betas, betaindex = [], {}   # τ-dependent factors β(τ)
    basis, bindex = [], {}      # basis function m = (beta index, integer power of k)

    def _split(term):
        """term -> (numeric coeff, τ-part, integer k power)"""
        coeff, rest = term.as_coeff_Mul()
        if k not in rest.free_symbols:
            return coeff, rest, 0
        indep, dep = rest.as_independent(k, as_Add=False)
        kp = dep.as_powers_dict().get(k, 0) if dep != 1 else 0
        if dep != k**kp:  # e.g. a/(b k² + c k²): cancel/factor first
            coeff, rest = sp.factor(sp.cancel(term)).as_coeff_Mul()
            indep, dep = rest.as_independent(k, as_Add=False)
            kp = dep.as_powers_dict().get(k, 0) if dep != 1 else 0
        assert dep == k**kp, f"non power-law k dependence in {term}"
        return coeff, indep, int(kp)

    def _basis(indep, kp):
        assert -3 <= kp <= 3
        if indep not in betaindex:
            betaindex[indep] = len(betas)
            betas.append(indep)
        key = (betaindex[indep], kp)
        if key not in bindex:
            bindex[key] = len(basis)
            basis.append(key)
        return bindex[key]

    def reg(e, whole=False):
        """entry expression -> list of (coeff, basis index); basis = k^e * beta(τ).
        whole=True keeps a sum as a single basis function (hub vectors p, q: one term per row).  Long sums (w0wa fluid rows)
        are grouped by power of k so that an entry costs at most one slot per distinct power."""
        if whole:
            c, indep, kp = _split(sp.factor(sp.cancel(e)))
            return [(float(c), _basis(indep, kp))]
        parts = [_split(t) for t in sp.Add.make_args(sp.expand(e))]
        if len(parts) <= 2:
            return [(float(c), _basis(indep, kp)) for c, indep, kp in parts]
        groups = {}
        for c, indep, kp in parts:
            groups[kp] = groups.get(kp, 0) + c * indep
        return [(1.0, _basis(sp.factor(g), kp)) for kp, g in sorted(groups.items())]

    reg(sp.Integer(1))  # basis 0 == 1
    rows = [[] for _ in range(N)]  # J_local CSR rows: (col, coeff, basis); duplicates (same col) allowed
    hub = dict(p=[], q=[], phi=[], psi=[])  # sparse term lists (index, coeff, basis)
    for i, fi in enumerate(m.f):
        for s, c in _linear_terms(fi, list(m.u) + [Phd, Psi]).items():
            terms = reg(c, whole=(s == Phd or s == Psi))
            if s == Phd:
                assert len(terms) == 1
                hub["p"] += [(i, cf, b) for cf, b in terms]
            elif s == Psi:
                assert len(terms) == 1
                hub["q"] += [(i, cf, b) for cf, b in terms]
            else:
                rows[i] += [(uidx[s], cf, b) for cf, b in terms]
        rows[i].sort()
    Phd_full = sp.expand(m.Phd_expr.subs(Psi, m.Psi_expr))
    for s, c in _linear_terms(Phd_full, m.u).items():
        hub["phi"] += [(uidx[s], cf, b) for cf, b in reg(c)]
    for s, c in _linear_terms(sp.expand(m.Psi_expr), m.u).items():
        hub["psi"] += [(uidx[s], cf, b) for cf, b in reg(c)]
    for v in hub.values():
        v.sort()
    L.rows, L.hub, L.basis, L.betas = rows, hub, basis, betas

    # ---- symbolic factorisation of the pattern of J_local: leaf peeling -> elimination forest + 2-core blocks
    adj = [set() for _ in range(N)]
    for i in range(N):
        for (j, _, _) in rows[i]:
            if j != i:
                adj[i].add(j)
                adj[j].add(i)
    alive = [True] * N
    deg = [len(a) for a in adj]
    parent, level = [-1] * N, [0] * N
    children = [[] for _ in range(N)]
    while True:  # rounds: peel all current leaves simultaneously (halves the depth of free chains)
        leaves = [i for i in range(N) if alive[i] and deg[i] == 1]
        if not leaves:
            break
        for i in leaves:
            if not (alive[i] and deg[i] == 1):
                continue  # its only neighbour was peeled in this round: i became a root
            j = next(v for v in adj[i] if alive[v])
            parent[i] = j
            children[j].append(i)
            level[j] = max(level[j], level[i] + 1)
            alive[i] = False
            deg[j] -= 1
    roots = [i for i in range(N) if alive[i]]
    blocks, seen = [], set()
    for r in roots:
        if r in seen:
            continue
        comp, stack = [], [r]
        seen.add(r)
        while stack:
            v = stack.pop()
            comp.append(v)
            for w in adj[v]:
                if alive[w] and w not in seen:
                    seen.add(w)
                    stack.append(w)
        blocks.append(sorted(comp))
    L.parent, L.level, L.children, L.blocks = parent, level, children, blocks
    L.nlevels = max(level) + 1
    blk_of = {}
    for b, comp in enumerate(blocks):
        for pos, v in enumerate(comp):
            blk_of[v] = (b, pos)
    L.blk_of = blk_of
    # roles of J_local entries: 0 diag, 1 "up" (row child, col parent), 2 "lo" (row parent, col child), 3 block off-diagonal
    roles = []
    for i in range(N):
        rr = []
        for (j, _, _) in rows[i]:
            if i == j:
                rr.append((0, i))
            elif parent[i] == j:
                rr.append((1, i))
            elif parent[j] == i:
                rr.append((2, j))
            else:
                assert i in blk_of and j in blk_of and blk_of[i][0] == blk_of[j][0], f"entry ({i},{j}) neither tree edge nor block"
                rr.append((3, blk_of[j][1]))
        roles.append(rr)
    L.roles = roles
    L.nnz_local = len(set((i, j) for i in range(N) for (j, _, _) in rows[i]))
    pat = set((i, j) for i in range(N) for (j, _, _) in rows[i])
    pi_, qi_ = {t[0] for t in hub["p"]}, {t[0] for t in hub["q"]}
    phj, psj = {t[0] for t in hub["phi"]}, {t[0] for t in hub["psi"]}
    pat |= {(i, j) for i in pi_ for j in phj} | {(i, j) for i in qi_ for j in psj} | {(i, i) for i in range(N)}
    L.nnz_full = len(pat)
    L.pattern = sorted(pat)
    # exact flop model of one f = J u product and one factorisation / solve with these tables
    nslots = sum(len(r) for r in rows)
    nhub = {n: len(v) for n, v in hub.items()}
    L.flops_f = 3 * nslots + 3 * (nhub["phi"] + nhub["psi"]) + 3 * (nhub["p"] + nhub["q"])
    nchild = sum(len(c) for c in children)
    blk_lu = sum(2 * len(b) ** 3 // 3 for b in blocks)
    blk_sv = sum(2 * len(b) ** 2 for b in blocks)
    L.flops_lu = 2 * nslots + 3 * nchild + blk_lu
    L.flops_solve = 2 * nchild + blk_sv + 3 * nchild
    return L


def _arr(ctype, name, vals, fmt="{}"):
    body = ", ".join(fmt.format(v) for v in vals) if len(vals) else "0"
    n = max(1, len(vals))
    return f"SB_TABLE {ctype} {name}[{n}] = {{{body}}};"


def generate(m: Model):
    """Returns (header_text, info dict)."""
    L = analyse(m)
    N, NB, NBETA = m.N, len(L.basis), len(L.betas)
    tau, k = m.tau, m.k
    S = m.S
    prog = m.prog.stmts
    ysyms = list(m.y)
    ypsyms = sp.symbols("yp_a yp_kap yp_XH yp_XHe yp_DT", real=True)
    info = dict(N=N, NB=NB, NBETA=NBETA, nnz_local=L.nnz_local, nnz_full=L.nnz_full, nlevels=L.nlevels, blocks=[len(b) for b in L.blocks],
                lmax=m.lmax, nx=m.nx, w0wa=m.w0wa, unames=m.unames, flops_f=L.flops_f, flops_lu=L.flops_lu, flops_solve=L.flops_solve)
    subs = {p: sp.Symbol(f"P[{i}]") for i, p in enumerate(m.par_list)}
    subs.update({s: sp.Symbol(f"y[{i}]") for i, s in enumerate(ysyms)})
    subs.update({s: sp.Symbol(f"yp[{i}]") for i, s in enumerate(ypsyms)})
    subs.update({s: sp.Symbol(f"u[{i}]") for i, s in enumerate(m.u)})
    info["npar"] = len(m.par_list)
    info["par_names"] = [str(p)[2:] for p in m.par_list]
    out = []
    W = out.append
    key = hashlib.sha1(f"v2-{m.lmax}-{m.nx}-{m.w0wa}".encode()).hexdigest()[:10]
    W("// GENERATED by symboltz.jl_b200/codegen/lower.py -- do not edit.")
    W(f"// model: lmax={m.lmax} nx={m.nx} w0wa={int(m.w0wa)}  N={N} NB={NB} NBETA={NBETA} nnz(J_local)={L.nnz_local} nnz(W)={L.nnz_full}")
    W("#pragma once")
    W("#include <math.h>")
    W("#ifndef SB_HD\n#ifdef __CUDACC__\n#define SB_HD __host__ __device__\n#else\n#define SB_HD\n#endif\n#endif")
    W("#ifndef SB_TABLE\n#ifdef __CUDACC__\n#define SB_TABLE static __device__ const\n#else\n#define SB_TABLE static const\n#endif\n#endif")
    nslots = sum(len(r) for r in L.rows)
    for name, val in [("SB_N", N), ("SB_NB", NB), ("SB_NBETA", NBETA), ("SB_NPAR", len(m.par_list)), ("SB_LMAX", m.lmax), ("SB_NX", m.nx), ("SB_W0WA", int(m.w0wa)),
                      ("SB_NSLOT", nslots), ("SB_NNZ_FULL", L.nnz_full), ("SB_NLEVELS", L.nlevels), ("SB_NBLOCKS", len(L.blocks)),
                      ("SB_MAXBLOCK", max(len(b) for b in L.blocks)), ("SB_BLKSTORE", sum(len(b) ** 2 for b in L.blocks)),
                      ("SB_P_KAPPA0", m.PAR_NAMES.index("kappa0")), ("SB_P_TAU0", m.PAR_NAMES.index("tau0")),
                      ("SB_FLOPS_F", L.flops_f), ("SB_FLOPS_LU", L.flops_lu), ("SB_FLOPS_SOLVE", L.flops_solve)]:
        W(f"#define {name} {val}")
    W(f'#define SB_MODEL_KEY "{key}"')
    for nm in ["Phi", "tb", "F0", "F2", "G0", "G2"]:
        W(f"#define SB_I_{nm.upper()} {m.unames.index(nm)}")

    # ---------------- tables
    rowptr, cols, coefs, bidx, rkind, ridx = [0], [], [], [], [], []
    for i in range(N):
        for (j, c, b), (rk, ri) in zip(L.rows[i], L.roles[i]):
            cols.append(j); coefs.append(c); bidx.append(b); rkind.append(rk); ridx.append(ri)
        rowptr.append(len(cols))
    W(_arr("short", "sb_rowptr", rowptr))
    W(_arr("short", "sb_col", cols))
    W(_arr("double", "sb_coef", coefs, "{!r}"))
    W(_arr("short", "sb_bidx", bidx))
    W(_arr("signed char", "sb_rkind", rkind))
    W(_arr("short", "sb_ridx", ridx))
    W(_arr("short", "sb_basis_beta", [b[0] for b in L.basis]))
    W(_arr("signed char", "sb_basis_kpow", [b[1] for b in L.basis]))
    hptr, hidx, hcoef, hb = [0], [], [], []
    for nm in ("p", "q", "phi", "psi"):
        for (i, c, b) in L.hub[nm]:
            hidx.append(i); hcoef.append(c); hb.append(b)
        hptr.append(len(hidx))
    W("// hub vectors p, q, phi, psi as sparse term lists: ranges [sb_hptr[v], sb_hptr[v+1]) for v = 0..3")
    W(_arr("short", "sb_hptr", hptr))
    W(_arr("short", "sb_hidx", hidx))
    W(_arr("double", "sb_hcoef", hcoef, "{!r}"))
    W(_arr("short", "sb_hb", hb))
    W(_arr("short", "sb_parent", L.parent))
    W(_arr("short", "sb_level", L.level))
    chptr, chidx = [0], []
    for i in range(N):
        chidx += L.children[i]
        chptr.append(len(chidx))
    W(_arr("short", "sb_chptr", chptr))
    W(_arr("short", "sb_chidx", chidx))
    order = sorted(range(N), key=lambda i: (L.level[i], i))
    lvptr = [0]
    for lv in range(L.nlevels):
        lvptr.append(lvptr[-1] + sum(1 for i in range(N) if L.level[i] == lv))
    W(_arr("short", "sb_lvorder", order))
    W(_arr("short", "sb_lvptr", lvptr))
    bptr, bvert, boff = [0], [], [0]
    for comp in L.blocks:
        bvert += comp
        bptr.append(len(bvert))
        boff.append(boff[-1] + len(comp) ** 2)
    W(_arr("short", "sb_bptr", bptr))
    W(_arr("short", "sb_bvert", bvert))
    W(_arr("short", "sb_boff", boff))
    blkid, blkpos = [-1] * N, [-1] * N
    for v, (b, pos) in L.blk_of.items():
        blkid[v], blkpos[v] = b, pos
    W(_arr("short", "sb_blkid", blkid))
    W(_arr("short", "sb_blkpos", blkpos))

    # ---------------- per-lane schedules for the integrator (warp = 32 lanes), layout [item][lane] (one coalesced line per item).
    # Path decomposition of the elimination forest: a path is a maximal chain v0 -> v1 -> ... in which every vertex after the
    # first has exactly one child (the previous one).  One lane owns one path and runs the elimination / substitution
    # recurrences along it in registers.  Paths starting at a leaf form phase 0.  What sits above them is the "top":
    #   * top blocks: the dense 2-core blocks, grown by every higher-phase path coupled to them (e.g. F1 joins {F2,G0,G1,G2}):
    #     inverted explicitly (pivoted Gauss-Jordan), applied as a mat-vec, lane = 8*block + row;
    #   * root paths: the remaining higher-phase paths; they must be roots of the forest fed by phase-0 paths only, so
    #     that their forward and backward substitutions complete in the same step as the block mat-vec.
    # A solve is then three warp-synchronous steps: phase-0 forward | top | phase-0 backward.
    # The integrator works in a RELABELLED state order in which every path is a contiguous index range (phase-0 paths, root
    # paths, top blocks), so that the recurrences address shared memory as base + compile-time offset; sb_nat / sb_newidx
    # convert at the kernel boundary (ICs, saved states).
    in_dense = {v for b in L.blocks if len(b) > 1 for v in b}
    path_of, paths, pphase = {}, [], []
    for v in sorted(range(N), key=lambda i: L.level[i]):
        if v in in_dense:
            continue
        ch = L.children[v]
        if len(ch) == 1:
            pid = path_of[ch[0]]
            paths[pid].append(v)
        else:
            pid = len(paths)
            paths.append([v])
            pphase.append(0 if not ch else 1 + max(pphase[path_of[c]] for c in ch))
        path_of[v] = pid
    nbr = [set() for _ in range(N)]
    for i in range(N):
        for (j, _, _) in L.rows[i]:
            if i != j:
                nbr[i].add(j)
                nbr[j].add(i)
    top_sets = [set(b) for b in L.blocks if len(b) > 1]
    absorbed, grew = set(), True
    while grew:
        grew = False
        for pid, p in enumerate(paths):
            if pphase[pid] == 0 or pid in absorbed:
                continue
            hit = [t for t in top_sets if any(w in t for v in p for w in nbr[v])]
            if hit:
                assert len(hit) == 1, "a path bridging two dense blocks is not supported"
                hit[0].update(p)
                absorbed.add(pid)
                grew = True
    top_blocks = [sorted(t) for t in top_sets]
    tblk_of = {v: (bi, pos) for bi, b in enumerate(top_blocks) for pos, v in enumerate(b)}
    toff = [0]
    for b in top_blocks:
        toff.append(toff[-1] + len(b) ** 2)
    TOPSTORE, TOPMAX = toff[-1], max([len(b) for b in top_blocks] + [1])
    paths0 = [p for pid, p in enumerate(paths) if pphase[pid] == 0]
    rpaths = [p for pid, p in enumerate(paths) if pphase[pid] > 0 and pid not in absorbed]
    for pid, p in enumerate(paths):
        if pphase[pid] > 0 and pid not in absorbed:
            assert pphase[pid] == 1 and L.parent[p[-1]] < 0, "higher-phase paths must be forest roots fed by leaf paths"
    for v in tblk_of:
        assert all(c not in tblk_of and pphase[path_of[c]] == 0 for c in L.children[v] if c not in tblk_of)
    assert len(top_blocks) <= 3 and all(len(b) <= 8 for b in top_blocks)
    PL = max(len(p) for p in paths0)
    PR = (len(paths0) + 31) // 32
    # single-vertex root paths share the top-block lanes' gather code (lanes after the block lanes) as long as they fit in the warp
    singles = [p[0] for p in rpaths if len(p) == 1][:max(0, 32 - 8 * len(top_blocks))]
    allpaths = paths0 + rpaths
    rpaths = [p for p in rpaths if not (len(p) == 1 and p[0] in singles)]
    TL = max([len(p) for p in rpaths] + [1])
    TR = (len(rpaths) + 31) // 32
    # new index -> natural index: top blocks first, then the paths by decreasing length of their longest J_local row, so that
    # the long rows share a round of the row-parallel sweeps (row i belongs to lane i % 32, round i / 32) and the ELL
    # storage of each round is only as wide as its longest row
    nat = []
    for b in top_blocks:
        nat += b
    ordered = sorted(allpaths, key=lambda p: -max(len(L.rows[v]) for v in p))
    # Bank-aware tie-break.  In the path sweeps lane = path and the lanes address shared memory at (end of path) − q, so two
    # paths whose ends are congruent modulo 16 doubles (= 32 banks) serialise every access they make together.  Paths with the
    # same key are interchangeable as far as the ELL widths go: permute them so that the ends (weight 2: both substitution
    # sweeps) and the starts (weight 1: the factorisation sweep) of the phase-0 paths collide as little as possible.
    p0ids = {id(p) for p in paths0}

    def _cost(seq):
        pos, ends, starts = len(nat), [], []
        for p in seq:
            if id(p) in p0ids:
                starts.append((pos % 16, len(p)))
                ends.append(((pos + len(p) - 1) % 16, len(p)))
            pos += len(p)
        c = 0
        for lst, w in ((ends, 2), (starts, 1)):
            for a in range(len(lst)):
                for b in range(a):
                    if lst[a][0] == lst[b][0]:
                        c += w * min(lst[a][1], lst[b][1])
        return c

    import random
    rng = random.Random(12345)
    keys = [-max(len(L.rows[v]) for v in p) for p in ordered]
    classes = [[i for i in range(len(ordered)) if keys[i] == k] for k in sorted(set(keys))]
    classes = [c for c in classes if len(c) > 1]
    best, bestc = list(ordered), _cost(ordered)
    if classes and bestc > 0:
        for restart in range(20):
            cur = list(ordered)
            if restart:
                for c in classes:
                    perm = c[:]
                    rng.shuffle(perm)
                    vals = [ordered[i] for i in perm]
                    for i, v in zip(c, vals):
                        cur[i] = v
            curc = _cost(cur)
            for it in range(4000):
                c = rng.choice(classes)
                i, j = rng.sample(c, 2)
                cur[i], cur[j] = cur[j], cur[i]
                nc = _cost(cur)
                if nc <= curc:
                    curc = nc
                else:
                    cur[i], cur[j] = cur[j], cur[i]
                if curc == 0:
                    break
            if curc < bestc:
                best, bestc = list(cur), curc
            if bestc == 0:
                break
    for p in best:
        nat += p
    assert sorted(nat) == list(range(N)) and N <= 1022
    # roles of the J_local entries in this schedule: 0 diag, 1 "up" (row child, col parent), 2 "lo" (row parent, col child),
    # 3 top-block entry (target = offset into the block store)
    sroles = []
    for i in range(N):
        rr = []
        for (j, _, _) in L.rows[i]:
            if i == j:
                rr.append((0, 0))
            elif i in tblk_of and j in tblk_of:
                assert tblk_of[i][0] == tblk_of[j][0]
                bb, nbk = tblk_of[i][0], len(top_blocks[tblk_of[i][0]])
                rr.append((3, N + toff[bb] + tblk_of[i][1] * nbk + tblk_of[j][1]))  # the block store follows mm[N] in shared memory
            elif L.parent[i] == j:
                rr.append((1, i))
            elif L.parent[j] == i:
                rr.append((2, j))
            else:
                raise AssertionError(f"entry ({i},{j}) neither tree edge nor top-block entry")
        sroles.append(rr)
    # index width of the packed schedules: 8 bits when everything fits (cheaper byte extraction on the device), else 10
    BITS = 8 if (N <= 254 and NB <= 255 and TOPSTORE <= 255) else 10
    HB = 8 if BITS == 8 else 16          # width of the two-field words
    KSH = 2 * BITS                       # role kind position in the ELL index word
    TSH = 20 if BITS == 8 else 22        # role target position
    inv = [0] * N
    for new_i, old_i in enumerate(nat):
        inv[old_i] = new_i
    R = (N + 31) // 32
    wdr = [max([len(L.rows[nat[ni]]) for ni in range(r * 32, min(N, r * 32 + 32))] + [1]) for r in range(R)]  # ELL width per round
    eoff = [sum(wdr[:r]) for r in range(R)]
    ELLN = sum(wdr)
    assert NB <= 1023 and N + TOPSTORE <= (4095 if BITS == 8 else 1023)
    ell_coef = [0.0] * (ELLN * 32)
    ell_idx = [0] * (ELLN * 32)
    for i in range(N):
        ni = inv[i]
        lane, r = ni % 32, ni // 32
        for w, ((j, c, b), (rk, ri)) in enumerate(zip(L.rows[i], sroles[i])):
            tgt = inv[ri] if rk == 2 else (ri if rk == 3 else 0)
            ell_coef[(eoff[r] + w) * 32 + lane] = c
            ell_idx[(eoff[r] + w) * 32 + lane] = inv[j] | (b << BITS) | (rk << KSH) | (tgt << TSH)
    pq_coef = [0.0] * (R * 2 * 32)
    pq_idx = [0] * (R * 32)
    for (i, c, b) in L.hub["p"]:
        ni = inv[i]
        pq_coef[((ni // 32) * 2 + 0) * 32 + ni % 32] = c
        pq_idx[(ni // 32) * 32 + ni % 32] |= b
    for (i, c, b) in L.hub["q"]:
        ni = inv[i]
        pq_coef[((ni // 32) * 2 + 1) * 32 + ni % 32] = c
        pq_idx[(ni // 32) * 32 + ni % 32] |= b << HB
    TPH = (len(L.hub["phi"]) + 31) // 32
    TPS = (len(L.hub["psi"]) + 31) // 32

    def _terms(lst, T):
        cf, ix = [0.0] * (T * 32), [0] * (T * 32)
        for t, (i, c, b) in enumerate(lst):
            cf[(t // 32) * 32 + t % 32] = c
            ix[(t // 32) * 32 + t % 32] = inv[i] | (b << HB)
        return cf, ix
    phc, phi_ = _terms(L.hub["phi"], TPH)
    psc, psi_ = _terms(L.hub["psi"], TPS)
    # path descriptors [round][lane]: head = start | len<<8 | parent_of_last<<16 (255: root); kids = children of the first vertex.
    # Root paths start on the first lane after the top-block lanes so that the two kinds of top work do not share lanes.
    NONE = (1 << BITS) - 1
    NOKIDS = NONE | (NONE << BITS) | (NONE << (2 * BITS))
    NOPAR = 255 if BITS == 8 else 4095
    LSH, PSH = (8, 16) if BITS == 8 else (12, 20)

    def _kids(v):
        kids = [inv[c] for c in L.children[v] if c not in tblk_of]
        assert len(kids) <= 3
        kids += [NONE] * (3 - len(kids))
        return kids[0] | (kids[1] << BITS) | (kids[2] << (2 * BITS))

    def _halfwarp_lanes(plist):
        """Lanes of the (at most 32) paths of one round.  64-bit shared-memory accesses are served per half-warp and the sweeps address
        (end of path) − q from every lane at once, so two paths of one half-warp whose ends (substitution sweeps, weight 2) or starts
        (factorisation sweep, weight 1) are congruent modulo 16 doubles cost an extra wavefront per access: split the paths over the two
        half-warps so that as few as possible collide (round 1 put them on lanes 0, 1, 2, ...: all in one half-warp, three colliding pairs
        at N = 82, +50 % wavefronts in the sweeps, profiles/integrate_r2.md).  Lanes without a path get a dummy head (length 0) whose
        unconditional loads fall into a bank none of the half-warp's paths uses."""
        import random
        n = len(plist)
        ends = [inv[p[-1]] % 16 for p in plist]
        starts = [inv[p[0]] % 16 for p in plist]
        lens = [len(p) for p in plist]

        def cost(side):
            c = 0
            for a in range(n):
                for b in range(a):
                    if side[a] == side[b]:
                        w = min(lens[a], lens[b])
                        c += (2 * w if ends[a] == ends[b] else 0) + (w if starts[a] == starts[b] else 0)
            return c
        rng = random.Random(4242)
        best, bestc = None, None
        for restart in range(30):
            side = [(q + restart) % 2 for q in range(n)] if restart < 2 else [rng.randrange(2) for _ in range(n)]
            while side.count(0) > 16 or side.count(1) > 16:
                i = rng.randrange(n)
                side[i] = 1 - side[i] if side.count(side[i]) > 16 else side[i]
            c = cost(side)
            for it in range(3000):
                if c == 0:
                    break
                i = rng.randrange(n)
                j = rng.randrange(n)
                trial = list(side)
                if rng.random() < 0.5:
                    trial[i] = 1 - trial[i]
                else:
                    trial[i], trial[j] = trial[j], trial[i]
                if trial.count(0) > 16 or trial.count(1) > 16:
                    continue
                tc = cost(trial)
                if tc <= c:
                    side, c = trial, tc
            if bestc is None or c < bestc:
                best, bestc = list(side), c
            if bestc == 0:
                break
        lanes, nxt = [0] * n, [0, 16]
        for q in range(n):
            lanes[q] = nxt[best[q]]
            nxt[best[q]] += 1
        dummies = []
        for h in (0, 1):
            used = {ends[q] for q in range(n) if best[q] == h}
            free = [r for r in range(16) if r not in used] or [0]
            # a dummy of length 0 at `start` reads like a path that ends at start − 1; keep every offset (down to −PL − 1) inside [0, N)
            start = next(i for i in range(PL + 2, N) if (i - 1) % 16 == free[len(free) // 2])
            dummies.append(start)
        return lanes, dummies

    def _paths(plist, rounds, lane0, spread=False):
        head, kid = [0] * (rounds * 32), [NOKIDS] * (rounds * 32)
        for rd in range(rounds):
            chunk = plist[rd * 32:(rd + 1) * 32]
            if spread and chunk:
                lanes, dummies = _halfwarp_lanes(chunk)
                for lane in range(32):
                    head[rd * 32 + lane] = dummies[lane // 16] | (0 << LSH) | (NOPAR << PSH)
            else:
                lanes = [(q + lane0) % 32 for q in range(len(chunk))]
            for p, lane in zip(chunk, lanes):
                start = inv[p[0]]
                assert [inv[v] for v in p] == list(range(start, start + len(p)))
                par = L.parent[p[-1]]
                par = NOPAR if par < 0 else inv[par]
                head[rd * 32 + lane] = start | (len(p) << LSH) | (par << PSH)
                kid[rd * 32 + lane] = _kids(p[0])
        return head, kid
    p_head, p_kids = _paths(paths0, PR, 0, spread=True)
    r_head, r_kids = _paths(rpaths, max(1, TR), 8 * len(top_blocks) + len(singles))
    t_kids, t_vert = [NOKIDS] * 32, [NONE] * 32
    for q, v in enumerate(singles):
        t_kids[8 * len(top_blocks) + q] = _kids(v)
        t_vert[8 * len(top_blocks) + q] = inv[v]
    for bi, b in enumerate(top_blocks):
        for pos, v in enumerate(b):
            t_kids[bi * 8 + pos] = _kids(v)
            t_vert[bi * 8 + pos] = inv[v]
        assert [inv[v] for v in b] == list(range(inv[b[0]], inv[b[0]] + len(b)))
    for name, val in [("SB_IDXBITS", BITS), ("SB_R", R), ("SB_ELLN", ELLN), ("SB_TPH", TPH), ("SB_TPS", TPS), ("SB_NTOP", len(top_blocks)), ("SB_TOPMAX", TOPMAX),
                      ("SB_TOPSTORE", max(1, TOPSTORE)), ("SB_PL", PL), ("SB_PR", PR), ("SB_TL", TL), ("SB_TR", TR), ("SB_NSINGLE", len(singles)),
                      ("SB_UNIQUE_TARGETS", int(all(len({(rk, ri) for (rk, ri) in sroles[i] if rk >= 2}) == sum(1 for (rk, ri) in sroles[i] if rk >= 2) for i in range(N))))]:
        W(f"#define {name} {val}")
    _sel = lambda vals: " : ".join(f"(r) == {r} ? {v}" for r, v in enumerate(vals[:-1])) + (" : " if len(vals) > 1 else "") + str(vals[-1])
    W(f"#define SB_WDR(r) ({_sel(wdr)})   // ELL width of round r")
    W(f"#define SB_EOFF(r) ({_sel(eoff)})  // first ELL slot of round r")
    W(_arr("short", "sb_nat", nat))
    W(_arr("short", "sb_newidx", inv))
    for nm in ["Phi", "tb", "F0", "F2", "G0", "G2"]:  # the same unknowns in the integrator's (path-contiguous) order: fused source evaluation
        W(f"#define SB_J_{nm.upper()} {inv[m.unames.index(nm)]}")
    W(_arr("double", "sb_ell_coef", ell_coef, "{!r}"))
    W(_arr("unsigned int", "sb_ell_idx", ell_idx, "{}u"))
    W(_arr("double", "sb_pq_coef", pq_coef, "{!r}"))
    W(_arr("unsigned int", "sb_pq_idx", pq_idx, "{}u"))
    W(_arr("double", "sb_phi_coef", phc, "{!r}"))
    W(_arr("unsigned int", "sb_phi_idx", phi_, "{}u"))
    W(_arr("double", "sb_psi_coef", psc, "{!r}"))
    W(_arr("unsigned int", "sb_psi_idx", psi_, "{}u"))
    W(_arr("unsigned int", "sb_path_head", p_head, "{}u"))
    W(_arr("unsigned int", "sb_path_kids", p_kids, "{}u"))
    W(_arr("unsigned int", "sb_root_head", r_head, "{}u"))
    W(_arr("unsigned int", "sb_root_kids", r_kids, "{}u"))
    W(_arr("unsigned int", "sb_top_kids", t_kids, "{}u"))
    W(_arr("unsigned int", "sb_top_vert", t_vert, "{}u"))
    W(_arr("int", "sb_top_n", [len(b) for b in top_blocks]))
    W(_arr("int", "sb_top_off", toff[:-1]))
    W(_arr("int", "sb_top_start", [inv[b[0]] for b in top_blocks]))
    W(_arr("unsigned int", "sb_basis_pack", [(b[0] | ((b[1] + 3) << HB)) for b in L.basis], "{}u"))

    flops = {}
    # ---------------- β_m(τ) and dβ_m/dτ given dy/dτ = yp
    seeds = {ys: yps for ys, yps in zip(ysyms, ypsyms)}
    seeds[tau] = sp.Integer(1)
    beta_stmts = [(sp.Symbol(f"beta_{i}", real=True), e) for i, e in enumerate(L.betas)]
    full = prog + beta_stmts
    tstm, dmap = tangent(full, seeds, "d_")
    code, n = _emit(_merge(full, tstm), [(f"beta[{i}]", s) for i, (s, _) in enumerate(beta_stmts)] + [(f"betad[{i}]", dmap[s]) for (s, _) in beta_stmts for i in [int(str(s)[5:])]], subs)
    flops["beta"] = n
    W("\n// τ-dependent factors β_m(τ) of the Jacobian entries (entry = const · k^e · β_m) and dβ_m/dτ for dy/dτ = yp")
    W("SB_HD static inline void sb_beta(double tau, const double* y, const double* yp, const double* P, double* beta, double* betad) {")
    W(code)
    W("}")

    # ---------------- initial conditions and Δm
    code, n = _emit(prog, [(f"u[{i}]", e) for i, e in enumerate(m.ic)], subs)
    flops["initial"] = n
    W("\n// closed-form adiabatic initial conditions at (τ,k)")
    W("SB_HD static inline void sb_initial(double tau, double k, const double* y, const double* P, double* u) {")
    W(code)
    W("}")
    code, n = _emit(prog, [("const double dm", m.Delta_m)], subs)
    flops["delta_m"] = n
    W("\n// total matter gauge-invariant overdensity Δm (c+b+h)")
    W("SB_HD static inline double sb_delta_m(double tau, double k, const double* y, const double* P, const double* u) {")
    W(code)
    W("    return dm;\n}")

    # ---------------- background RHS, Jacobian, κ derivatives
    g = list(m.g)
    code, n = _emit(prog, [(f"g[{i}]", e) for i, e in enumerate(g)], subs)
    flops["bg_rhs"] = n
    W("\n// background + thermodynamics right-hand side g(y), y = (a, _κ, XH⁺, XHe⁺, ΔT)")
    W("SB_HD static inline void sb_bg_rhs(const double* y, const double* P, double* g) {")
    W(code)
    W("}")
    jstm, jouts = [], []
    for j in range(5):
        tj, dj = tangent(prog, {ysyms[j]: sp.Integer(1)}, f"j{j}_")
        jstm.append(tj)
        jouts += [(f"J[{i * 5 + j}]", dj[g[i]]) for i in range(5)]
    code, n = _emit(_merge(prog, *jstm), [(f"g[{i}]", e) for i, e in enumerate(g)] + jouts, subs)
    flops["bg_jac"] = n
    W("// g and its analytic Jacobian (row-major 5x5)")
    W("SB_HD static inline void sb_bg_rhs_jac(const double* y, const double* P, double* g, double* J) {")
    W(code)
    W("}")
    # κ̈ = D κ̇, κ⃛ = D κ̈ with D = Σ_i g_i ∂/∂y_i (flow derivative; MTK expands D(κ̇) through the RHS in the same way)
    fseeds = {ys: gi for ys, gi in zip(ysyms, g)}
    t1, d1 = tangent(prog, fseeds, "f1_")
    p1 = _merge(prog, t1)
    t2, d2 = tangent(p1, fseeds, "f2_")
    code, n = _emit(_merge(p1, t2), [(f"g[{i}]", e) for i, e in enumerate(g)] + [("out[0]", g[1]), ("out[1]", d1[g[1]]), ("out[2]", d2[d1[g[1]]])], subs)
    flops["kappa_derivs"] = n
    W("\n// κ̇, κ̈, κ⃛ along the background flow (visibility function and its derivatives) and g(y)")
    W("SB_HD static inline void sb_kappa_derivs(const double* y, const double* P, double* g, double* out) {")
    W(code)
    W("}")
    obs_names = ["a", "Hc", "kd", "cs2", "Xe", "Tb"]
    code, _ = _emit(prog, [(f"o[{i}]", S[nm]) for i, nm in enumerate(obs_names)], subs)
    W("\n// diagnostics: a, ℋ, κ̇, c_s², X_e, T_b from y")
    W("SB_HD static inline void sb_bg_observe(const double* y, const double* P, double* o) {")
    W(code)
    W("}")
    info["flops"] = flops
    info["key"] = key
    info["L"] = L
    return "\n".join(out) + "\n", info
