"""Independent pin of the multi-layer urban solvers (test infrastructure).

A GLOBAL two-point boundary-value solve of one (column, spectral interval) of an
urban / vegetated-urban tile in arbitrary precision (mpmath), sharing NOTHING
with the oracle restatement (oracle/src) or the CUDA path beyond the inputs:

 * no eigen-decomposition, no reflectance / transmittance matrices, no adding
   method, no Schur inverse: every layer is represented by the matrix
   exponential of its full (2n+d)-order rate matrix
       d/dz [u; v; s] = G [u; v; s] (+ [-b; b; 0] in the longwave),
       G = [[-G1, -G2, -G3], [G2, G1, G3], [0, 0, G0]],  z downward from the layer top
   (radtool/radtool_calc_matrices_sw_eig.F90:158-166), the unknowns are the flux
   vectors at the top of every layer, and the interface conditions (overlap
   matrices, roofs, ground, incoming flux at canopy top) close ONE banded linear
   system for the whole column, solved by Gaussian elimination with partial
   pivoting;
 * layer-integrated fluxes come from the dense inverse of G applied to the
   boundary values, int x dz = G^-1 (x_base - x_top - c dz);
 * the geometry (region fractions, overlap, perimeters, exchange and wall
   rates) and the flux partition are written from the Fortran
   (radsurf/radsurf_urban_sw.F90:264-600, :603-984; radsurf_urban_lw.F90:296-546,
   :548-858; radsurf_overlap.F90:178-394; radsurf_norm_perim.F90:131-281) in
   plain Python loops.

With mp.dps = 150 the growing exponentials of thick layers are harmless.  Scope:
tile types 2 (urban) and 3 (vegetated urban), nreg <= 3, building fraction not
increasing with height and trees not overhanging buildings (asserted).
Returns the same per-field numbers the solvers write, for one column.
"""
import mpmath as mp

# the propagator of a layer holds exp(+lambda dz) and exp(-lambda dz) side by side; lambda dz reaches ~70
# in the fixtures (60 orders of magnitude each way), so 150 digits leave > 80 for the result
mp.mp.dps = 150
PI = mp.mpf("3.14159265358979323846")  # radiation_constants.F90:24 (the FP64 constant, as all builds use it)
F = mp.mpf


def _lg(lg):
    ns = int(lg.nstream)
    g = lambda a: [F(float(a[i])) for i in range(ns)]
    return dict(ns=ns, mu=g(lg.mu), tan=g(lg.tan_ang), w=g(lg.weight), hw=g(lg.hweight), vw=g(lg.vweight),
                vadj=F(float(lg.vadjustment)), vadj2=F(float(lg.vadjustment2)))


def _overlap_urban(nreg, fu, fl):
    """calc_overlap_matrix_max_ran_urban, no-overhang branches (radsurf_overlap.F90:178-280)."""
    O = [[F(0)] * (nreg + 1) for _ in range(nreg)]
    if nreg == 1:
        O[0][0], O[0][1] = fl[0], fl[1]
    elif nreg == 2:
        pc = max(fu[1], fl[1])
        assert pc <= fl[0] + fl[1], "trees overhanging buildings: outside the scope of this pin"
        O[0][2] = fl[2]
        O[0][0] = fl[0] + fl[1] - pc
        O[0][1] = pc - fu[1]
        O[1][0] = pc - fl[1]
        O[1][1] = fu[1] + fl[1] - pc
    else:
        pc = max(fu[1] + fu[2], fl[1] + fl[2])
        assert pc <= fl[0] + fl[1] + fl[2], "trees overhanging buildings: outside the scope of this pin"
        O[0][3] = fl[3]
        O[0][0] = fl[0] + fl[1] + fl[2] - pc
        if pc > fu[1] + fu[2]:
            O[1][1], O[2][2] = fu[1], fu[2]
            O[0][1], O[0][2] = fl[1] - fu[1], fl[2] - fu[2]
        else:
            O[1][1], O[2][2] = fl[1], fl[2]
            O[1][0], O[2][0] = fu[1] - fl[1], fu[2] - fl[2]
    return O


def _geometry(cfg, nreg, L, bf, bs, vf, vs, vcf):
    """Region fractions, directional overlap matrices U[k], V[k] for interfaces k = 0..L
    (k = base of layer k, 0-based; k = L canopy top), perimeters."""
    minv, minb = F(cfg.min_vegetation_fraction), F(cfg.min_building_fraction)
    frac = []
    for j in range(L):
        f0 = 1 - bf[j]
        if nreg > 1:
            f0 = max(F(0), f0 - vf[j])
            fv = max(F(0), 1 - bf[j] - f0) / (nreg - 1)
            frac.append([f0] + [fv] * (nreg - 1))
        else:
            frac.append([f0])
    frac.append([F(1)] + [F(0)] * (nreg - 1))
    U, V = [], []
    fl = [F(0)] * nreg + [sum(frac[0])]
    for k in range(L + 1):
        fu = frac[k]
        if k == 0:
            # ground interface: its U, V are never used (the ground albedo sets a_above directly,
            # urban_sw:591-602)
            O = [[F(0)] * (nreg + 1) for _ in range(nreg)]
        else:
            O = _overlap_urban(nreg, fu, fl)
        U.append([[O[a][b] / fl[b] if fl[b] >= minv else F(0) for b in range(nreg + 1)] for a in range(nreg)])
        V.append([[O[a][b] / fu[a] if fu[a] >= minv else F(0) for a in range(nreg)] for b in range(nreg + 1)])
        fl = list(fu) + [F(0)]
        if k < L - 1:
            fl[nreg] = sum(frac[k + 1]) - sum(frac[k])
            assert fl[nreg] >= 0, "building fraction increasing with height: outside the scope of this pin"
        elif k == L - 1:
            fl[nreg] = 1 - sum(frac[k])
    iso = F(cfg.vegetation_isolation_factor_urban)
    npm = [[F(0)] * nreg for _ in range(L)]
    npw = [[F(0)] * nreg for _ in range(L)]
    for j in range(L):
        if nreg > 1 and vf[j] > minv:
            den = max(minb, 1 - bf[j]) * vs[j]
            if cfg.use_symmetric_vegetation_scale_urban:
                p1 = 4 * vf[j] * max(F(0), 1 - vf[j] - bf[j]) / den
            else:
                p1 = 4 * vf[j] / vs[j]
            npm[j][0] = p1
            if nreg > 2:
                npm[j][nreg - 1] = iso / 2 * p1
                npm[j][0] = (1 - iso / 2) * p1
                if cfg.use_symmetric_vegetation_scale_urban:
                    npm[j][1] = (1 - iso) * 4 * (vf[j] / 2) * (1 - vf[j] / 2 - bf[j]) / den
                else:
                    npm[j][1] = (1 - iso) * 4 * vf[j] / (mp.sqrt(2) * vs[j])
        if bf[j] > minb:
            w = 4 * bf[j] / bs[j]
            npw[j][0] = w
            if nreg > 1:
                if 1 - vf[j] - bf[j] <= minv:
                    if nreg == 2:
                        npw[j][1] = w
                    else:
                        npw[j][1], npw[j][2] = w * (1 - iso), w * iso
                    npw[j][0] = F(0)
                elif vf[j] > minv and vcf[j] > 0:
                    if nreg == 2:
                        npw[j][1] = w * vcf[j]
                    else:
                        npw[j][1], npw[j][2] = w * vcf[j] * (1 - iso), w * vcf[j] * iso
                    npw[j][0] = w * (1 - vcf[j])
    return frac, U, V, npm, npw


def _rates(cfg, nreg, frac_j, npm_j, npw_j, wall_scale):
    minv = F(cfg.min_vegetation_fraction)
    fex = [[F(0)] * nreg for _ in range(nreg)]  # [to][from]
    for r in range(nreg - 1):
        if not (frac_j[r] <= minv or frac_j[r + 1] <= minv):
            fex[r + 1][r] = npm_j[r] / (PI * frac_j[r])
            fex[r][r + 1] = npm_j[r] / (PI * frac_j[r + 1])
    if nreg > 2 and npm_j[nreg - 1] > 0:
        if not (frac_j[2] <= minv or frac_j[0] <= minv):
            # the reference reads norm_perim(jreg) with the loop-exit value jreg = nreg (urban_sw:390-391)
            fex[0][2] = npm_j[nreg - 1] / (PI * frac_j[2])
            fex[2][0] = npm_j[nreg - 1] / (PI * frac_j[0])
    fw = [F(0) if frac_j[r] <= minv else npw_j[r] * wall_scale / (PI * frac_j[r]) for r in range(nreg)]
    return fex, fw


def _gammas(lg, regs, fex, fw, ext, ssa, wall_ext, wall_factor, tan0, sin0, zcos, sw):
    """Gamma matrices of the solved regions `regs` (urban_sw:426-494, urban_lw:394-445)."""
    ns, nr = lg["ns"], len(regs)
    n = nr * ns
    g0 = mp.zeros(nr, nr)
    g1 = mp.zeros(n, n)
    g2 = mp.zeros(n, n)
    g3 = mp.zeros(n, nr)
    for a, rf in enumerate(regs):
        for b, rt in enumerate(regs):
            if rf != rt:
                if sw:
                    g0[b, a] = tan0 * fex[rt][rf]
                for js in range(ns):
                    g1[js + b * ns, js + a * ns] = lg["tan"][js] * fex[rt][rf]
        # losses go to EVERY other region of the layer, solved or not (the diagonal is built before
        # the sub-block is cut out: urban_sw:426-441, 512-583)
        loss = sum(fex[rt][rf] for rt in range(len(fex)) if rt != rf)
        if sw:
            g0[a, a] = -tan0 * loss - ext[rf] / zcos - tan0 * fw[rf] * wall_ext
        for js in range(ns):
            i = js + a * ns
            g1[i, i] = -lg["tan"][js] * loss - ext[rf] / lg["mu"][js] - lg["tan"][js] * fw[rf] * wall_ext
            for jt in range(ns):
                g2[jt + a * ns, i] = (lg["w"][jt] * ext[rf] * ssa[rf] / lg["mu"][js]
                                      + lg["vw"][jt] * lg["tan"][js] * fw[rf] * wall_factor) / 2
            if sw:
                g3[i, a] = (lg["w"][js] * ext[rf] * ssa[rf] + lg["vw"][js] * sin0 * fw[rf] * wall_factor) / 2
    return g0, g1 + g2, g2, g3


class _Sparse:
    """Row-sparse linear system with Gaussian elimination and partial pivoting."""

    def __init__(self, n, nrhs):
        self.rows = [dict() for _ in range(n)]
        self.rhs = [[F(0)] * nrhs for _ in range(n)]
        self.n, self.nrhs = n, nrhs

    def add(self, i, j, v):
        if v != 0:
            self.rows[i][j] = self.rows[i].get(j, F(0)) + v

    def solve(self):
        n, rows, rhs = self.n, self.rows, self.rhs
        cols = [set() for _ in range(n)]
        for i, r in enumerate(rows):
            for j in r:
                cols[j].add(i)
        done = [False] * n
        order = []
        for k in range(n):
            cand = [i for i in cols[k] if not done[i] and rows[i].get(k, 0) != 0]
            p = max(cand, key=lambda i: abs(rows[i][k]))
            done[p] = True
            order.append(p)
            piv = rows[p][k]
            for i in cand:
                if i == p:
                    continue
                l = rows[i][k] / piv
                del rows[i][k]
                for j, v in rows[p].items():
                    if j != k:
                        rows[i][j] = rows[i].get(j, F(0)) - l * v
                        cols[j].add(i)
                for c in range(self.nrhs):
                    rhs[i][c] -= l * rhs[p][c]
        x = [[F(0)] * self.nrhs for _ in range(n)]
        for k in reversed(range(n)):
            p = order[k]
            for c in range(self.nrhs):
                s = rhs[p][c]
                for j, v in rows[p].items():
                    if j != k:
                        s -= v * x[j][c]
                x[k][c] = s / rows[p][k]
        return x


def _sic_sweeps(env):
    """The reference's upward and downward recurrences with the independent layer operators."""
    g_ = lambda k: env[k]
    sw, L, n, d, ns, nreg, N = g_("sw"), g_("L"), g_("n"), g_("d"), g_("ns"), g_("nreg"), g_("N")
    layers, U, V, hw, zcos, spec, spl, bf, frac = (g_(k) for k in
                                                   ("layers", "U", "V", "hw", "zcos", "spec", "spl", "bf", "frac"))
    icol, g, cfg = g_("icol"), g_("g"), g_("cfg")
    m, nrb = n + ns, nreg + 1
    Z = mp.zeros
    ops = []
    for j in range(L):
        ly = layers[j]
        regs, nn, dd, M, p, G = ly["regs"], ly["nn"], ly["dd"], ly["M"], ly["p"], ly["G"]
        Muu, Muv, Mvu, Mvv = M[:nn, :nn], M[:nn, nn:2 * nn], M[nn:2 * nn, :nn], M[nn:2 * nn, nn:2 * nn]
        Mi = mp.inverse(Muu)
        Rs = -Mi * Muv
        Ts = Mvv + Mvu * Rs
        Gi = mp.inverse(G)
        K = Gi[:nn, :] + Gi[nn:2 * nn, :]

        def emb(A, rows_s, cols_s, nr_, nc_, rk, ck):
            """sub-block -> full size; rk/ck: entries per region (ns for streams, 1 for beams)."""
            Fm = Z(nr_, nc_)
            for a, ra in enumerate(regs):
                for ia in range(rk):
                    for b, rb in enumerate(regs):
                        for ib in range(ck):
                            Fm[ra * rk + ia, rb * ck + ib] = A[a * rk + ia, b * ck + ib]
            return Fm

        def embv(v):
            w = Z(n, 1)
            for a, ra in enumerate(regs):
                for js in range(ns):
                    w[ra * ns + js] = v[a * ns + js]
            return w

        o = dict(R=emb(Rs, nn, nn, n, n, ns, ns), T=emb(Ts, nn, nn, n, n, ns, ns))
        if sw:
            Mus, Mvs, Mss = M[:nn, 2 * nn:], M[nn:2 * nn, 2 * nn:], M[2 * nn:, 2 * nn:]
            Sup = -Mi * Mus
            Sdn = Mvs + Mvu * Sup
            o.update(Sup=emb(Sup, nn, dd, n, d, ns, 1), Sdn=emb(Sdn, nn, dd, n, d, ns, 1),
                     E=emb(Mss, dd, dd, d, d, 1, 1), Idiff=emb(K[:, :nn], nn, nn, n, n, ns, ns),
                     Idd=emb(-K[:, 2 * nn:], nn, dd, n, d, ns, 1),
                     Idir=emb(-Gi[2 * nn:, 2 * nn:], dd, dd, d, d, 1, 1))
        else:
            src = -Mi * p[:nn, 0]
            I_ = mp.eye(nn)
            IF = -K[:, :nn] * Rs + K[:, nn:2 * nn] * (Ts - I_)
            isrc = -K[:, :nn] * src + K[:, nn:2 * nn] * src - K * ly["c"] * env["dz"][j]
            o.update(src=embv(src), IF=emb(IF, nn, nn, n, n, ns, ns), isrc=embv(isrc))
        ops.append(o)

    def UxI(Uk):  # (U (x) I_ns): n x m
        A = Z(n, m)
        for u in range(nreg):
            for lo in range(nrb):
                for js in range(ns):
                    A[u * ns + js, lo * ns + js] = Uk[u][lo]
        return A

    def VxI(Vk):  # m x n
        A = Z(m, n)
        for lo in range(nrb):
            for up in range(nreg):
                for js in range(ns):
                    A[lo * ns + js, up * ns + js] = Vk[lo][up]
        return A

    Vm = lambda Vk: mp.matrix([[Vk[lo][up] for up in range(nreg)] for lo in range(nrb)])
    a_above, d_above, s_above = [Z(n, n)], [Z(n, max(d, 1))], [Z(n, 1)]
    a_below, d_below, s_below, Dinv = [None], [None], [None], []
    if sw:
        galb = F(float(spec.ground_albedo[icol, g]))
        galb_dir = F(float((spec.ground_albedo_dir if cfg.use_sw_direct_albedo else spec.ground_albedo)[icol, g]))
    else:
        gemis, gemission = F(float(spec.ground_emissivity[icol, g])), F(float(spec.ground_emission[icol, g]))
    for r in range(nreg):
        for jt in range(ns):
            for jf in range(ns):
                a_above[0][r * ns + jt, r * ns + jf] = (galb if sw else 1 - gemis) * hw[jt]
            if sw:
                d_above[0][r * ns + jt, r] = zcos * galb_dir * hw[jt]
            else:
                s_above[0][r * ns + jt] = hw[jt] * frac[0][r] * gemission
    I_n = mp.eye(n)
    for j in range(L):
        o = ops[j]
        Di = mp.inverse(I_n - a_above[j] * o["R"])
        Dinv.append(Di)
        ab = Z(m, m)
        ab[:n, :n] = o["R"] + o["T"] * Di * a_above[j] * o["T"]
        if sw:
            ralb = spl(spec.roof_albedo, j)
            ralb_dir = spl(spec.roof_albedo_dir, j) if spec.roof_albedo_dir is not None else ralb
            db = Z(m, d + 1)
            db[:n, :d] = o["Sup"] + o["T"] * Di * (d_above[j] * o["E"] + a_above[j] * o["Sdn"])
        else:
            ralb = 1 - spl(spec.roof_emissivity, j)
            exposed = max(F(0), bf[j] - bf[j + 1]) if j < L - 1 else bf[j]
            sb = Z(m, 1)
            sb[:n, 0] = o["src"] + o["T"] * Di * (s_above[j] + a_above[j] * o["src"])
        for js in range(ns):
            for jf in range(ns):
                ab[n + js, n + jf] = ralb * hw[js]
            if sw:
                db[n + js, d] = zcos * ralb_dir * hw[js]
            else:
                sb[n + js] = hw[js] * spl(spec.roof_emission, j) * exposed
        a_below.append(ab)
        Uj, Vj = UxI(U[j + 1]), VxI(V[j + 1])
        a_above.append(Uj * ab * Vj)
        if sw:
            d_below.append(db)
            d_above.append(Uj * db * Vm(V[j + 1]))
        else:
            s_below.append(sb)
            s_above.append(Uj * sb)
    res = dict(top=[[None] * L, [None] * L], base=[[None] * L, [None] * L], idiff=[[None] * L, [None] * L],
               idir=[[None] * L, [None] * L], up_top=[None, None])
    aL = a_above[L]
    hwv = mp.matrix([hw[js] for js in range(ns)])
    if sw:
        res["up_top"][0] = sum(d_above[L][js, 0] for js in range(ns)) / zcos
        res["up_top"][1] = sum((aL[:ns, :ns] * hwv)[js] for js in range(ns))
    else:
        res["up_top"][0] = sum(s_above[L][js] for js in range(ns))
        res["up_top"][1] = sum((aL[:ns, :ns] * hwv)[js] for js in range(ns))
    for ps in range(2):
        direct, emit = sw and ps == 0, (not sw) and ps == 0
        dn_above = Z(n, 1)
        dir_above = Z(max(d, 1), 1)
        if direct:
            dir_above[0] = 1 / zcos
        if ps == 1:
            for js in range(ns):
                dn_above[js] = hw[js]
        for j in reversed(range(L)):
            o = ops[j]
            dn_below = VxI(V[j + 1]) * dn_above
            up_below = a_below[j + 1] * dn_below
            if sw:
                dir_below = Vm(V[j + 1]) * dir_above if direct else Z(d + 1, 1)
                up_below = up_below + d_below[j + 1] * dir_below
                dir_new = o["E"] * dir_below[:d, 0]
                refl = d_above[j] * dir_new
                dn_new = Dinv[j] * (o["T"] * dn_below[:n, 0] + o["R"] * refl + o["Sdn"] * dir_below[:d, 0])
                up_new = a_above[j] * dn_new + refl
                cv = dn_below[:n, 0] - dn_new - up_below[:n, 0] + up_new
                ddir = dir_below[:d, 0] - dir_new
                idf = o["Idiff"] * cv + o["Idd"] * ddir
                idr = o["Idir"] * ddir
            else:
                if emit:
                    up_below = up_below + s_below[j + 1]
                    dn_new = Dinv[j] * (o["T"] * dn_below[:n, 0] + o["R"] * s_above[j] + o["src"])
                    up_new = a_above[j] * dn_new + s_above[j]
                else:
                    dn_new = Dinv[j] * (o["T"] * dn_below[:n, 0])
                    up_new = a_above[j] * dn_new
                idf = o["IF"] * (dn_below[:n, 0] + up_new) + (o["isrc"] if emit else Z(n, 1))
                dir_below, dir_new, idr = Z(1, 1), Z(1, 1), None
            res["top"][ps][j] = [up_below[i] for i in range(n)] + [dn_below[i] for i in range(n)] + (
                [dir_below[i] for i in range(d)] if sw else [])
            res["base"][ps][j] = [up_new[i] for i in range(n)] + [dn_new[i] for i in range(n)] + (
                [dir_new[i] for i in range(d)] if sw else [])
            res["idiff"][ps][j] = [idf[i] for i in range(n)]
            res["idir"][ps][j] = [idr[i] for i in range(d)] if sw else []
            dn_above, dir_above = dn_new, dir_new
    return res


def solve_column(cfg, cp, spec, icol, g, band, mode="bvp"):
    """One column (0-based) and spectral interval of an urban / vegetated-urban tile.
    band = "sw": returns {"sw_norm_dir": {...}, "sw_norm_diff": {...}, "bc": {...}};
    band = "lw": {"lw_internal": ..., "lw_norm": ..., "bc": ...}.  Per-layer fields are lists
    over the layers of the column (ground upwards).

    mode = "bvp": the global boundary-value solve described above (the physically consistent
    solution).  mode = "sic": the same independent layer operators (matrix exponential, dense
    inverse of G) pushed through the reference's OWN sweep recurrences
    (radsurf_urban_sw.F90:603-984, radsurf_urban_lw.F90:548-858), which differ from the
    boundary-value solution for more than one stream: the downward pass solves
    (I - a_above R) x = ... where flux continuity requires (I - R a_above) (see
    tests/test_bvp_pin.py); the two agree on everything the upward sweep alone determines
    (top-of-canopy albedo / emissivity / emission, top_net) and on every field for one stream."""
    sw = band == "sw"
    tile = int(cp.i_representation[icol])
    assert tile in (2, 3)
    nreg = 1 if tile == 2 else int(cfg.n_vegetation_region_urban) + 1
    lg = _lg(cfg.lg_sw_urban if sw else cfg.lg_lw_urban)
    ns = lg["ns"]
    n, d = nreg * ns, (nreg if sw else 0)
    N = 2 * n + d
    L, il1 = int(cp.nlay[icol]), int(cp.istartlay[icol]) - 1
    lay = lambda a, j: F(float(a[il1 + j]))
    spl = lambda a, j: F(float(a[il1 + j, g]))
    dz = [lay(cp.dz, j) for j in range(L)]
    bf = [lay(cp.building_fraction, j) for j in range(L)]
    bs = [lay(cp.building_scale, j) for j in range(L)]
    veg = nreg > 1
    vf = [lay(cp.veg_fraction, j) if veg else F(0) for j in range(L)]
    vs = [lay(cp.veg_scale, j) if veg else F(1) for j in range(L)]
    ve = [lay(cp.veg_ext, j) if veg else F(0) for j in range(L)]
    vcf = [lay(cp.veg_contact_fraction, j) if veg else F(0) for j in range(L)]
    minv, minb = F(cfg.min_vegetation_fraction), F(cfg.min_building_fraction)
    frac, U, V, npm, npw = _geometry(cfg, nreg, L, bf, bs, vf, vs, vcf)
    if sw:
        cos_sza = F(float(cp.cos_sza[icol]))
        zcos = max(cos_sza, F("1e-6"))
        sin0 = mp.sqrt(1 - zcos * zcos)
        tan0 = sin0 / zcos
    else:
        zcos = sin0 = tan0 = F(0)
    hw = lg["hw"]

    # ---- per layer: solved regions, rate matrix, propagator, inverse --------------------------
    layers = []
    for j in range(L):
        od = [F(1)] * nreg
        if nreg == 3:
            fsd = lay(cp.veg_fsd, j)
            od[1] = mp.exp(-fsd * (1 + fsd / 2 * (1 + fsd / 2)))
            od[2] = 2 - od[1]
        air_ext, air_ssa = spl(spec.air_ext, j), spl(spec.air_ssa, j)
        vssa = spl(spec.veg_ssa, j) if veg else F(0)
        ext, ssa = [air_ext], [air_ssa]
        for r in range(1, nreg):
            e = air_ext + od[r] * ve[j]
            ext.append(e)
            ssa.append((air_ext * air_ssa + od[r] * ve[j] * vssa) / max(e, F("1e-8")))
        fex, fw = _rates(cfg, nreg, frac[j], npm[j], npw[j], F(1) if sw else lg["vadj2"])
        if sw:
            walb, wspec = spl(spec.wall_albedo, j), spl(spec.wall_specular_frac, j)
            wall_ext, wall_factor = 1 - walb * wspec, walb * (1 - wspec)
        else:
            # sic: interval 1 for every interval (urban_lw:392)
            wall_ext, wall_factor = F(1), 1 - F(float(spec.wall_emissivity[il1 + j, 0]))
        if veg and vf[j] <= minv:
            regs = [0]
        elif veg and frac[j][0] <= minv:
            regs = list(range(1, nreg))
        else:
            regs = list(range(nreg))
        g0, g1, g2, g3 = _gammas(lg, regs, fex, fw, ext, ssa, wall_ext, wall_factor, tan0, sin0, zcos, sw)
        nr = len(regs)
        nn, dd = nr * ns, (nr if sw else 0)
        G = mp.zeros(2 * nn + dd, 2 * nn + dd)
        for a in range(nn):
            for b in range(nn):
                G[a, b], G[a, nn + b] = -g1[a, b], -g2[a, b]
                G[nn + a, b], G[nn + a, nn + b] = g2[a, b], g1[a, b]
            for b in range(dd):
                G[a, 2 * nn + b], G[nn + a, 2 * nn + b] = -g3[a, b], g3[a, b]
        for a in range(dd):
            for b in range(dd):
                G[2 * nn + a, 2 * nn + b] = g0[a, b]
        c = mp.zeros(2 * nn + dd, 1)
        book = {}
        if not sw:
            planck = [spl(spec.clear_air_planck, j)]
            vpl = spl(spec.veg_planck, j) if veg else F(0)
            vapl = spl(spec.veg_air_planck, j) if veg else F(0)
            for r in range(1, nreg):
                planck.append((air_ext * (1 - air_ssa) * vapl + od[r] * ve[j] * (1 - vssa) * vpl)
                              / max(ext[r] * (1 - ssa[r]), F("1e-8")))
            wem = spl(spec.wall_emission, j)
            efac = 2 * sum(hw[js] / lg["mu"][js] for js in range(ns))
            vol = [frac[j][r] * (ext[r] * (1 - ssa[r]) * planck[r]) for r in range(nreg)]
            for a, r in enumerate(regs):
                for js in range(ns):
                    b = hw[js] / lg["mu"][js] * vol[r] + lg["vw"][js] / 2 * (npw[j][r] * lg["vadj"] * wem)
                    c[a * ns + js] = -b
                    c[nn + a * ns + js] = b
            book = dict(reg=[efac * v for v in vol],
                        air=[F(0)] + [efac * frac[j][r] * air_ext * (1 - air_ssa) * vapl for r in range(1, nreg)],
                        veg=[F(0)] + [efac * frac[j][r] * ve[j] * (1 - vssa) * vpl * od[r] for r in range(1, nreg)],
                        wall=sum(npw[j]) * lg["vadj"] * wem)
        # propagator of the augmented system [x; 1]: x_base = M x_top + p
        A = mp.zeros(2 * nn + dd + 1, 2 * nn + dd + 1)
        for a in range(2 * nn + dd):
            for b in range(2 * nn + dd):
                A[a, b] = G[a, b] * dz[j]
            A[a, 2 * nn + dd] = c[a] * dz[j]
        Ex = mp.expm(A)
        M = Ex[:2 * nn + dd, :2 * nn + dd]
        p = Ex[:2 * nn + dd, 2 * nn + dd]
        layers.append(dict(regs=regs, nn=nn, dd=dd, M=M, p=p, G=G, c=c, od=od, fw=fw, ext=ext, ssa=ssa,
                           air_ext=air_ext, air_ssa=air_ssa, vssa=vssa, book=book))

    sic = None
    if mode == "sic":
        sic = _sic_sweeps(locals())
    else:
        # ---- global system: unknown x_top of every layer in FULL-size indexing --------------------
        # x = [u (n) ; v (n) ; s (d)] per layer; entries of unsolved regions are pinned to zero at the
        # layer's outgoing sides (the reference zeroes R, T, S of those regions: urban_sw:512-583).
        nrhs = 2
        S = _Sparse(N * L, nrhs)
        iu = lambda j, i: N * j + i
        iv = lambda j, i: N * j + n + i
        isd = lambda j, r: N * j + 2 * n + r

        def base_row(j, kind, i):
            """Coefficients of (x_base of layer j)[kind, i] on the unknowns: {col: coef}, const."""
            ly = layers[j]
            regs, nn, dd = ly["regs"], ly["nn"], ly["dd"]
            r, js = (i // ns, i % ns) if kind != "s" else (i, 0)
            if r not in regs:
                return {}, F(0)
            a = regs.index(r)
            row = {"u": a * ns + js, "v": nn + a * ns + js, "s": 2 * nn + a}[kind]
            co = {}
            for b, rb in enumerate(regs):
                for jb in range(ns):
                    co[iu(j, rb * ns + jb)] = ly["M"][row, b * ns + jb]
                    co[iv(j, rb * ns + jb)] = ly["M"][row, nn + b * ns + jb]
                if dd:
                    co[isd(j, rb)] = ly["M"][row, 2 * nn + b]
            return co, ly["p"][row]

        if sw:
            galb = F(float(spec.ground_albedo[icol, g]))
            galb_dir = F(float((spec.ground_albedo_dir if cfg.use_sw_direct_albedo else spec.ground_albedo)[icol, g]))
        else:
            gemis = F(float(spec.ground_emissivity[icol, g]))
            gemission = F(float(spec.ground_emission[icol, g]))
        for j in range(L):
            ly = layers[j]
            regs = ly["regs"]
            # (1) downward fluxes at the top of layer j come through interface j+1 from the base of layer j+1
            Vk = V[j + 1]
            for r in range(nreg):
                for js in range(ns):
                    row = iv(j, r * ns + js)
                    S.add(row, row, F(1))
                    for up in range(nreg):
                        if Vk[r][up] == 0:
                            continue
                        if j == L - 1:
                            # incoming at canopy top: RHS 0 = direct / internal pass, RHS 1 = diffuse / normalised
                            if up == 0:
                                S.rhs[row][1] += Vk[r][up] * hw[js]
                        else:
                            co, const = base_row(j + 1, "v", up * ns + js)
                            for col, v in co.items():
                                S.add(row, col, -Vk[r][up] * v)
                            S.rhs[row][0] += Vk[r][up] * const  # emission only in the internal pass
                if sw:
                    row = isd(j, r)
                    S.add(row, row, F(1))
                    for up in range(nreg):
                        if Vk[r][up] == 0:
                            continue
                        if j == L - 1:
                            if up == 0:
                                S.rhs[row][0] += Vk[r][up] / zcos
                        else:
                            co, const = base_row(j + 1, "s", up)
                            for col, v in co.items():
                                S.add(row, col, -Vk[r][up] * v)
            # (2) upward fluxes at the base of layer j: ground, or interface j from the top of layer j-1
            for r in range(nreg):
                for js in range(ns):
                    i = r * ns + js
                    row = iu(j, i)
                    if r not in regs:
                        S.add(row, row, F(1))  # unsolved region: nothing leaves its top
                        continue
                    co, const = base_row(j, "u", i)
                    for col, v in co.items():
                        S.add(row, col, v)
                    S.rhs[row][0] -= const
                    if j == 0:
                        # u_base = albedo * hweight * sum_js v_base (same region) [+ direct / emission]
                        for jf in range(ns):
                            co2, c2 = base_row(0, "v", r * ns + jf)
                            refl = galb if sw else (1 - gemis)
                            for col, v in co2.items():
                                S.add(row, col, -refl * hw[js] * v)
                            S.rhs[row][0] += refl * hw[js] * c2
                        if sw:
                            co2, _ = base_row(0, "s", r)
                            for col, v in co2.items():
                                S.add(row, col, -zcos * galb_dir * hw[js] * v)
                        else:
                            S.rhs[row][0] += hw[js] * frac[0][r] * gemission
                    else:
                        Uk = U[j]
                        for lo in range(nreg):
                            if Uk[r][lo] != 0:
                                S.add(row, iu(j - 1, lo * ns + js), -Uk[r][lo])
                        if Uk[r][nreg] != 0:
                            # roof of layer j-1: reflects what interface j sends down onto it
                            Vr = V[j][nreg]
                            if sw:
                                ralb = spl(spec.roof_albedo, j - 1)
                                ralb_dir = spl(spec.roof_albedo_dir, j - 1) if spec.roof_albedo_dir is not None else ralb
                            else:
                                ralb = 1 - spl(spec.roof_emissivity, j - 1)
                                exposed = max(F(0), bf[j - 1] - bf[j])
                                S.rhs[row][0] += Uk[r][nreg] * hw[js] * spl(spec.roof_emission, j - 1) * exposed
                            for up in range(nreg):
                                if Vr[up] == 0:
                                    continue
                                for jf in range(ns):
                                    co2, c2 = base_row(j, "v", up * ns + jf)
                                    for col, v in co2.items():
                                        S.add(row, col, -Uk[r][nreg] * ralb * hw[js] * Vr[up] * v)
                                    S.rhs[row][0] += Uk[r][nreg] * ralb * hw[js] * Vr[up] * c2
                                if sw:
                                    co2, _ = base_row(j, "s", up)
                                    for col, v in co2.items():
                                        S.add(row, col, -Uk[r][nreg] * zcos * ralb_dir * hw[js] * Vr[up] * v)
        X = S.solve()

    # ---- fluxes and partition --------------------------------------------------------------
    names = ("sw_norm_dir", "sw_norm_diff") if sw else ("lw_internal", "lw_norm")
    out = {nm: {} for nm in names}
    out["bc"] = {}
    lay_fields = ["roof_in", "roof_net", "wall_in", "wall_net", "clear_air_abs", "veg_abs", "veg_air_abs",
                  "flux_dn_layer_top", "flux_up_layer_top", "flux_dn_layer_base", "flux_up_layer_base"]
    if sw:
        lay_fields += ["roof_in_dir", "wall_in_dir", "veg_abs_dir", "flux_dn_dir_layer_top", "flux_dn_dir_layer_base"]
    for ps, nm in enumerate(names):
        o = out[nm]
        for k in lay_fields:
            o[k] = [F(0)] * L
        emit = (not sw) and ps == 0
        direct = sw and ps == 0
        xb_above = None  # fluxes at the base of the layer above (full size)
        for j in reversed(range(L)):
            ly = layers[j]
            regs, nn, dd = ly["regs"], ly["nn"], ly["dd"]
            if sic is not None:
                top, base = sic["top"][ps][j], sic["base"][ps][j]
            else:
                top = [X[N * j + i][ps] for i in range(N)]
                base = [F(0)] * N
                for kind, off, cnt in (("u", 0, n), ("v", n, n), ("s", 2 * n, d)):
                    for i in range(cnt):
                        co, const = base_row(j, kind, i)
                        base[off + i] = sum(v * X[col][ps] for col, v in co.items()) + (const if emit else 0)
            # what interface j+1 sends down (incl. onto the roof of layer j)
            if j == L - 1:
                v_ab = [F(0)] * n
                s_ab = [F(0)] * nreg
                if direct:
                    s_ab[0] = 1 / zcos
                if ps == 1:
                    for js in range(ns):
                        v_ab[js] = hw[js]
            else:
                v_ab, s_ab = xb_above[n:2 * n], (xb_above[2 * n:] if sw else [F(0)] * nreg)
            Vr = V[j + 1][nreg]
            roof_diff = sum(Vr[up] * v_ab[up * ns + js] for up in range(nreg) for js in range(ns))
            roof_dir = sum(Vr[up] * s_ab[up] for up in range(nreg)) if sw else F(0)
            if sw:
                ralb = spl(spec.roof_albedo, j)
                ralb_dir = spl(spec.roof_albedo_dir, j) if spec.roof_albedo_dir is not None else ralb
                roof_up = ralb * roof_diff + (zcos * ralb_dir * roof_dir if direct else 0)
                if direct:
                    o["roof_in_dir"][j] = zcos * roof_dir
                o["roof_in"][j] = (zcos * roof_dir if direct else 0) + roof_diff
            else:
                exposed = max(F(0), bf[j] - bf[j + 1]) if j < L - 1 else bf[j]
                roof_up = (1 - spl(spec.roof_emissivity, j)) * roof_diff + (
                    spl(spec.roof_emission, j) * exposed if emit else 0)
                o["roof_in"][j] = roof_diff
            o["roof_net"][j] = o["roof_in"][j] - roof_up
            # integrated fluxes of the solved block: G^-1 (x_base - x_top - c dz)
            sel = [a * ns + js for a in range(len(regs)) for js in range(ns)]
            full = lambda off, a, js: off + regs[a] * ns + js
            rhs = mp.zeros(2 * nn + dd, 1)
            for a in range(len(regs)):
                for js in range(ns):
                    rhs[a * ns + js] = base[full(0, a, js)] - top[full(0, a, js)]
                    rhs[nn + a * ns + js] = base[full(n, a, js)] - top[full(n, a, js)]
                if dd:
                    rhs[2 * nn + a] = base[2 * n + regs[a]] - top[2 * n + regs[a]]
            if emit:
                rhs = rhs - ly["c"] * dz[j]
            idiff = [[F(0)] * ns for _ in range(nreg)]
            idir = [F(0)] * nreg
            if sic is not None:
                for r in range(nreg):
                    for js in range(ns):
                        idiff[r][js] = sic["idiff"][ps][j][r * ns + js]
                    if d:
                        idir[r] = sic["idir"][ps][j][r]
            else:
                integ = mp.lu_solve(ly["G"], rhs)
                for a, r in enumerate(regs):
                    for js in range(ns):
                        idiff[r][js] = integ[a * ns + js] + integ[nn + a * ns + js]
                    if dd:
                        idir[r] = integ[2 * nn + a]
            smu = [sum(idiff[r][js] / lg["mu"][js] for js in range(ns)) for r in range(nreg)]
            stan = [sum(idiff[r][js] * lg["tan"][js] for js in range(ns)) for r in range(nreg)]
            air_abs = ly["air_ext"] * (1 - ly["air_ssa"])
            vabs = ve[j] * (1 - ly["vssa"])
            bk = ly["book"]
            o["clear_air_abs"][j] = air_abs * ((idir[0] if direct else 0) + smu[0]) - (bk["reg"][0] * dz[j] if emit else 0)
            for r in range(1, nreg):
                o["veg_air_abs"][j] += air_abs * ((idir[r] if direct else 0) + smu[r]) - (bk["air"][r] * dz[j] if emit else 0)
                o["veg_abs"][j] += vabs * ((idir[r] if direct else 0) + smu[r]) * ly["od"][r] - (
                    bk["veg"][r] * dz[j] if emit else 0)
                if direct:
                    o["veg_abs_dir"][j] += vabs * idir[r] * ly["od"][r]
            win_dir = sum(ly["fw"][r] * sin0 * idir[r] for r in range(nreg)) if direct else F(0)
            win = win_dir + sum(ly["fw"][r] * stan[r] for r in range(nreg))
            o["wall_in"][j] = win
            if sw:
                if direct:
                    o["wall_in_dir"][j] = win_dir
                o["wall_net"][j] = win * (1 - spl(spec.wall_albedo, j))
            else:
                o["wall_net"][j] = win * spl(spec.wall_emissivity, j) - (bk["wall"] * dz[j] if emit else 0)
            sdir_t = zcos * sum(top[2 * n:]) if direct else F(0)
            sdir_b = zcos * sum(base[2 * n:]) if direct else F(0)
            if sw:
                o["flux_dn_dir_layer_top"][j], o["flux_dn_dir_layer_base"][j] = sdir_t, sdir_b
            o["flux_dn_layer_top"][j] = sdir_t + sum(top[n:2 * n])
            o["flux_up_layer_top"][j] = sum(top[:n])
            o["flux_dn_layer_base"][j] = sdir_b + sum(base[n:2 * n])
            o["flux_up_layer_base"][j] = sum(base[:n])
            xb_above = base
        base0 = xb_above
        gdir = zcos * sum(base0[2 * n:]) if direct else F(0)
        o["ground_dn"] = gdir + sum(base0[n:2 * n])
        o["ground_net"] = o["ground_dn"] - sum(base0[:n])
        o["ground_vertical_diff"] = sum((base0[n + r * ns + js] + base0[r * ns + js]) * lg["tan"][js] / PI
                                        for r in range(nreg) for js in range(ns))
        if sw:
            o["ground_dn_dir"] = gdir
            o["top_dn_dir"] = F(1) if direct else F(0)
        # flux leaving the canopy top: interface L maps the top of layer L-1 (and its roof) upwards
        topL = sic["top"][ps][L - 1] if sic is not None else [X[N * (L - 1) + i][ps] for i in range(N)]
        up_top = F(0)
        Uk = U[L]
        for r in range(nreg):
            for js in range(ns):
                s = sum(Uk[r][lo] * topL[lo * ns + js] for lo in range(nreg))
                # roof of the top layer
                if Uk[r][nreg] != 0:
                    jtop = L - 1
                    Vr = V[L][nreg]
                    inc_diff = Vr[0] * sum(hw) if ps == 1 else F(0)
                    inc_dir = Vr[0] / zcos if direct else F(0)
                    if sw:
                        ralb = spl(spec.roof_albedo, jtop)
                        ralb_dir = spl(spec.roof_albedo_dir, jtop) if spec.roof_albedo_dir is not None else ralb
                        ru = ralb * hw[js] * inc_diff + zcos * ralb_dir * hw[js] * inc_dir
                    else:
                        ru = (1 - spl(spec.roof_emissivity, jtop)) * hw[js] * inc_diff + (
                            hw[js] * spl(spec.roof_emission, jtop) * bf[jtop] if emit else 0)
                    s += Uk[r][nreg] * ru
                up_top += s
        if sic is not None:
            up_top = sic["up_top"][ps]  # from the upward sweep, as the reference does (urban_sw:672-674)
        o["top_dn"] = F(0) if emit else F(1)
        o["top_net"] = o["top_dn"] - up_top
        if sw:
            out["bc"]["sw_albedo_dir" if direct else "sw_albedo"] = up_top
        elif emit:
            out["bc"]["lw_emission"] = up_top
        else:
            out["bc"]["lw_emissivity"] = 1 - up_top
    return out
