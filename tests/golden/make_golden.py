"""Generates the golden fixtures under tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
It imports the reference's own folding/utils_ros/utils_ros.py with a stub
``pyrosetta`` module (the only thing gen_rst needs from it is the import line)
and the reference's utils_trX2dy/utils.py geometry with stub Bio/matplotlib.
Nothing under tests/, bench.py or smoke() reads /root/reference at run time.

Outputs
  example_{NMR,Xray}.npz, example_seq.fasta   the reference's example inputs (data fixtures)
  example_natives_ca.npz                       CA traces of example/apo.pdb, holo.pdb
  gen_rst_example_NMR.npz                      a,b,p per type + sha256 of all text lines + every 7th table
  gen_rst_random24.npz                         full tables for a random L=24 input (edge cases: p near cutoffs)
  geometry_random.npz                          reference get_dihedrals/get_angles on random points
  gen_rst_variants24.npz                       gen_idp_rst / gen_gpcr_rst / gen_rst_af2 tables on random L=24 inputs
  example_tmscore.npz                          bin/TMscore (TM-score, RMSD) + GloCon on the 8 example decoys and 2 natives
  example_backbone_stats.npz                   bond / angle statistics of the reference's 8 example decoys (Rosetta-written)
  dynamics_example48.npz                       outer-loop arithmetic (get_neighbors, pros, process_distribution...)
  example_reliability.npz                      backbones of the 8 example decoys + the reference's ramachandran_score of each
"""
import hashlib, os, shutil, sys, tempfile, types
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_gen_rst():
    sys.modules["pyrosetta"] = types.ModuleType("pyrosetta")
    sys.path.insert(0, os.path.join(REF, "folding"))
    from utils_ros import utils_ros  # noqa
    return utils_ros


def load_reference_geometry():
    for name in ("Bio", "Bio.PDB", "matplotlib", "matplotlib.pyplot"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["Bio.PDB"].PDBParser = object
    sys.modules["Bio.PDB"].PPBuilder = object
    sys.modules["Bio"].PDB = sys.modules["Bio.PDB"]
    sys.path.insert(0, REF)
    import importlib
    return importlib.import_module("utils_trX2dy.utils")


def run_gen_rst(utils_ros, npz, seq, use_orient=True):
    import json
    params = json.load(open(os.path.join(REF, "folding/data/params.json")))
    params["USE_ORIENT"] = use_orient
    params["seq"] = seq
    tmp = tempfile.TemporaryDirectory(prefix="/dev/shm/")
    rst = utils_ros.gen_rst(npz, tmp, params)
    res = {}
    for name, recs in rst.items():
        a = np.array([r[0] for r in recs], dtype=np.int32)
        b = np.array([r[1] for r in recs], dtype=np.int32)
        p = np.array([r[2] for r in recs], dtype=np.float32)
        lines, texts = [], []
        for r in recs:
            toks = r[3].split()
            fn = [t for t in toks if t.startswith(tmp.name)][0]
            texts.append(open(fn).read())
            lines.append(r[3].replace(tmp.name, "TMP"))
        res[name] = dict(a=a, b=b, p=p, texts=texts, lines=lines)
    tmp.cleanup()
    return res


def run_variant(utils_ros, fn_name, npz, seq, use_orient, known=None):
    """gen_idp_rst / gen_rst_af2 / gen_gpcr_rst of the reference (utils_ros.py:148-654), same packing as run_gen_rst."""
    import json
    params = json.load(open(os.path.join(REF, "folding/data/params.json")))
    params["USE_ORIENT"] = use_orient
    params["seq"] = seq
    tmp = tempfile.TemporaryDirectory(prefix="/dev/shm/")
    fn = getattr(utils_ros, fn_name)
    rst = fn(npz, known, tmp, params) if known is not None else fn(npz, tmp, params)
    res = {}
    for name, recs in rst.items():
        if name == "rep":
            continue
        a = np.array([r[0] for r in recs], dtype=np.int32)
        b = np.array([r[1] for r in recs], dtype=np.int32)
        p = np.array([r[2] for r in recs], dtype=np.float32)
        lines, texts = [], []
        for r in recs:
            toks = r[3].split()
            f = [t for t in toks if t.startswith(tmp.name)][0]
            texts.append(open(f).read())
            lines.append(r[3].replace(tmp.name, "TMP"))
        res[name] = dict(a=a, b=b, p=p, texts=texts, lines=lines)
    tmp.cleanup()
    return res


def make_variant_golden():
    """Restraint variants of SURVEY 8a row 15 on a random L=24 input: -r idp (idr mask picks the
    max-bin energy reference), -r af2 (64-bin CA-CA distogram, 60 knots), -r gpcr (template-blended
    tables), plus add_idr_rst's selection for mode 3."""
    ur = load_reference_gen_rst()
    g = np.load(f"{HERE}/gen_rst_random24.npz")
    rnd = {k: g[f"in_{k}"] for k in ("dist", "omega", "theta", "phi")}
    L = rnd["dist"].shape[0]
    rng = np.random.default_rng(2424)
    idr = rng.random((L, L)) < 0.3
    idr = idr | idr.T
    inp = dict(rnd, idr=idr)
    out = {"in_idr": idr}
    for orient in (True, False):
        tag = "idp" if orient else "idp_noorient"
        for k, v in pack(run_variant(ur, "gen_idp_rst", inp, "A" * L, orient), every=1).items():
            out[f"{tag}__{k}"] = v
    # gpcr: 5 templates given as real-valued maps (distance in A, angles in rad)
    M = 5
    base = np.abs(rng.normal(size=(L, L))) * 6 + 3.0
    base = 0.5 * (base + base.T)
    known = dict(dist=np.stack([base + rng.normal(size=(L, L)) * 1.0 for _ in range(M)]).astype(np.float64),
                 omega=rng.uniform(-np.pi, np.pi, size=(M, L, L)),
                 theta_asym=rng.uniform(-np.pi, np.pi, size=(M, L, L)),
                 phi_asym=rng.uniform(0, np.pi, size=(M, L, L)))
    known["dist"][:, rng.random((L, L)) < 0.2] = 25.0      # beyond the last bin -> 'no contact'
    for k, v in known.items():
        out[f"in_known_{k}"] = v
    for orient in (True, False):
        tag = "gpcr" if orient else "gpcr_noorient"
        for k, v in pack(run_variant(ur, "gen_gpcr_rst", inp, "A" * L, orient, known=known), every=1).items():
            out[f"{tag}__{k}"] = v
    # af2: 64-bin distogram over CA-CA, bin edges linspace(2.3125, 21.6875, 63)
    a = rng.dirichlet(np.full(64, 0.3), size=(L, L)).astype(np.float32)
    w = rng.uniform(0, 1, size=(L, L, 1)).astype(np.float32) ** 2
    e_last = np.zeros(64, dtype=np.float32); e_last[-1] = 1
    d64 = (w * a + (1 - w) * e_last).astype(np.float32)
    d64 = (0.5 * (d64 + d64.transpose(1, 0, 2))).astype(np.float32)
    bins = np.linspace(2.3125, 21.6875, 63)
    out["in_af2_dist"], out["in_af2_bins"] = d64, bins
    for k, v in pack(run_variant(ur, "gen_rst_af2", dict(dist=d64, bins=bins), "A" * L, False), every=1).items():
        out[f"af2__{k}"] = v
    np.savez_compressed(f"{HERE}/gen_rst_variants24.npz", **out)
    print("variant golden written")


def pack(res, every=1):
    out = {}
    for name, r in res.items():
        out[f"{name}_a"], out[f"{name}_b"], out[f"{name}_p"] = r["a"], r["b"], r["p"]
        h = hashlib.sha256()
        for t in r["texts"]:
            h.update(t.encode())
        out[f"{name}_sha256"] = np.array(h.hexdigest())
        hl = hashlib.sha256()
        for t in r["lines"]:
            hl.update(t.encode())
        out[f"{name}_lines_sha256"] = np.array(hl.hexdigest())
        idx = np.arange(0, len(r["texts"]), every)
        out[f"{name}_sub_idx"] = idx
        out[f"{name}_sub_text"] = np.array([r["texts"][k] for k in idx])
        out[f"{name}_line0"] = np.array(r["lines"][0] if r["lines"] else "")
    return out


def ca_trace(pdb):
    xyz = []
    for ln in open(pdb):
        if ln.startswith("ATOM") and ln[12:16].strip() == "CA":
            xyz.append([float(ln[30:38]), float(ln[38:46]), float(ln[46:54])])
    return np.array(xyz)


def main():
    ur = load_reference_gen_rst()
    seq = "".join(l.strip() for l in open(f"{REF}/example/seq.fasta") if not l.startswith(">"))
    for tag in ("NMR", "Xray"):
        shutil.copy(f"{REF}/example/output/seq/pred_npz/seq_{tag}.npz", f"{HERE}/example_{tag}.npz")
    shutil.copy(f"{REF}/example/seq.fasta", f"{HERE}/example_seq.fasta")
    np.savez_compressed(f"{HERE}/example_natives_ca.npz", apo=ca_trace(f"{REF}/example/apo.pdb"),
                        holo=ca_trace(f"{REF}/example/holo.pdb"))

    npz = np.load(f"{HERE}/example_NMR.npz")
    np.savez_compressed(f"{HERE}/gen_rst_example_NMR.npz", **pack(run_gen_rst(ur, npz, seq), every=7))

    # random small input: Dirichlet rows, some rows pushed to the 'no contact' bin so
    # that probabilities straddle the 0.05 / 0.55 / 0.65 cut-offs
    rng = np.random.default_rng(24)
    L = 24
    def rand(nb, sym):
        a = rng.dirichlet(np.full(nb, 0.3), size=(L, L)).astype(np.float32)
        w = rng.uniform(0, 1, size=(L, L, 1)).astype(np.float32) ** 2
        e0 = np.zeros(nb, dtype=np.float32); e0[0] = 1
        a = (w * a + (1 - w) * e0).astype(np.float32)
        if sym:
            a = (0.5 * (a + a.transpose(1, 0, 2))).astype(np.float32)
        return a
    rnd = dict(dist=rand(37, True), omega=rand(25, True), theta=rand(25, False), phi=rand(13, False))
    packed = pack(run_gen_rst(ur, rnd, "A" * L), every=1)
    packed.update({f"in_{k}": v for k, v in rnd.items()})
    np.savez_compressed(f"{HERE}/gen_rst_random24.npz", **packed)
    packed = pack(run_gen_rst(ur, rnd, "A" * L, use_orient=False), every=1)
    np.savez_compressed(f"{HERE}/gen_rst_random24_noorient.npz", **packed)

    # geometry: the reference's numpy dihedral / angle on random points
    ug = load_reference_geometry()
    pts = rng.normal(size=(4, 64, 3)) * 5.0
    np.savez_compressed(f"{HERE}/geometry_random.npz", pts=pts,
                        dihedral=ug.get_dihedrals(pts[0].copy(), pts[1].copy(), pts[2].copy(), pts[3].copy()),
                        angle=ug.get_angles(pts[0].copy(), pts[1].copy(), pts[2].copy()))
    print("golden fixtures written to", HERE)




def make_dynamics_golden():
    """Reference outer-loop arithmetic (get_neighbors, pros, process_distribution...) on the first
    48 residues of the example (a reference decoy + the NMR distograms)."""
    ug = load_reference_geometry()
    seq = "".join(l.strip() for l in open(f"{REF}/example/seq.fasta") if not l.startswith(">"))[:48]
    at = {k: [] for k in ("N", "CA", "C", "CB")}
    res = {}
    for ln in open(f"{REF}/example/output/seq/pred_pdb/conf_1_1.pdb"):
        if ln.startswith("ATOM") and ln[12:16].strip() in at:
            res.setdefault(int(ln[22:26]), {})[ln[12:16].strip()] = [float(ln[30:38]), float(ln[38:46]), float(ln[46:54])]
    ids = sorted(res)[:48]
    xyzs = {"N": np.array([res[i]["N"] for i in ids]), "CA": np.array([res[i]["CA"] for i in ids]),
            "C": np.array([res[i]["C"] for i in ids]), "CB": {k: res[i]["CB"] for k, i in enumerate(ids) if "CB" in res[i]}}
    cb = np.array([res[i].get("CB", [np.nan] * 3) for i in ids])
    key, d6, o6, t6, p6 = ug.get_neighbors(xyzs, seq, 20)
    assert key is False
    fact = ug.pros(d6[None], o6[None], t6[None], p6[None], angle=True)
    fd, ft, fo, fp = (fact[k][0, 0] for k in range(4))           # order: dist, theta, omega, phi
    npz = np.load(f"{HERE}/example_NMR.npz")
    crop = {k: np.ascontiguousarray(npz[k][:48, :48]) for k in ("dist", "omega", "theta", "phi")}
    out = dict(seq=np.array(seq), n=xyzs["N"], ca=xyzs["CA"], c=xyzs["C"], cb=cb, d6=d6, o6=o6, t6=t6, p6=p6,
               jd=fd.argmax(-1), jo=fo.argmax(-1), jt=ft.argmax(-1), jp=fp.argmax(-1))
    for k, f in (("dist", fd), ("omega", fo), ("theta", ft), ("phi", fp)):
        out[f"in_{k}"] = crop[k]
        out[f"proc_{k}"] = ug.process_distribution_with_pred_distribution(crop[k], f, norm=True, smooth=True, sigma=1.0)
    out["tmp"] = ug.process_distribution_with_pred_distribution(crop["dist"], fd, norm=False)
    np.savez_compressed(f"{HERE}/dynamics_example48.npz", **out)
    print("dynamics golden written")


def make_tmscore_golden():
    """The reference's structure-comparison step (utils_trX2dy/utils.py:514-540) shells out to the
    prebuilt bin/TMscore.  Run it here on the reference's 8 example decoys + the two natives (all
    45 pairs, both normalisations) and keep TM-score / RMSD next to the CA and CB traces."""
    import itertools, re, subprocess
    exe = "/tmp/TMscore_ref"
    shutil.copy(f"{REF}/bin/TMscore", exe)
    os.chmod(exe, 0o755)
    names = ["apo", "holo"] + [f"conf_{a}_{b}" for a in (1, 2) for b in (1, 2, 3, 4)]
    paths = [f"{REF}/example/{n}.pdb" if n in ("apo", "holo") else f"{REF}/example/output/seq/pred_pdb/{n}.pdb" for n in names]
    ug = load_reference_geometry()

    def backbone(pdb):
        at = {}
        for ln in open(pdb):
            if ln.startswith("ATOM") and ln[12:16].strip() in ("N", "CA", "C", "CB"):
                at.setdefault(int(ln[22:26]), {})[ln[12:16].strip()] = [float(ln[30:38]), float(ln[38:46]), float(ln[46:54])]
        ids = sorted(at)
        n, ca, c = (np.array([at[i][k] for i in ids]) for k in ("N", "CA", "C"))
        b, cc = ca - n, c - ca
        cb = -0.58273431 * np.cross(b, cc) + 0.56802827 * b - 0.54067466 * cc + ca      # virtual CB (utils.py:132-135)
        for k, i in enumerate(ids):
            if "CB" in at[i]:
                cb[k] = at[i]["CB"]
        return ca, cb, n

    ca, cb, nn = zip(*[backbone(p) for p in paths])
    M = len(names)
    tm, rm = np.zeros((M, M)), np.zeros((M, M))
    for i, j in itertools.product(range(M), repeat=2):
        if i == j:
            continue
        out = subprocess.run([exe, paths[i], paths[j]], stdout=subprocess.PIPE, universal_newlines=True).stdout
        rm[i, j] = float(re.search(r"RMSD of  the common residues=\s+([\d.]+)", out).group(1))
        tm[i, j] = float(re.search(r"TM-score    =\s+([\d.]+)", out).group(1))       # normalised by structure j
    # GloCon of the same set, by the reference's own arithmetic (utils.py:543-567) on get_neighbors' dist maps
    seq = "".join(l.strip() for l in open(f"{REF}/example/seq.fasta") if not l.startswith(">"))
    L = len(seq)
    glocon = np.zeros((M, M))
    dmaps = []
    cfull = []
    for p_ in paths:   # C atoms for the reference's own get_neighbors (utils_trX2dy/utils.py:125-182)
        at = {}
        for ln in open(p_):
            if ln.startswith("ATOM") and ln[12:16].strip() == "C":
                at[int(ln[22:26])] = [float(ln[30:38]), float(ln[38:46]), float(ln[46:54])]
        cfull.append(np.array([at[i] for i in sorted(at)]))
    for k in range(M):
        xyzs = {"N": nn[k], "CA": ca[k], "C": cfull[k], "CB": {i: cb[k][i] for i in range(L) if seq[i] != "G"}}
        key, d, _, _, _ = ug.get_neighbors(xyzs, seq, 20)          # the reference's distance map: 0 beyond 20 A
        assert key is False
        dmaps.append(d)
    for i, j in itertools.product(range(M), repeat=2):
        if i <= j:
            continue
        diff = np.abs(dmaps[i] - dmaps[j])
        diff[diff <= 3] = 0
        glocon[i, j] = np.sum(np.triu(diff)) / (len(diff) * (len(diff) - 1) / 2)
    glocon = glocon + glocon.T
    np.savez_compressed(f"{HERE}/example_tmscore.npz", names=np.array(names), ca=np.stack(ca), cb=np.stack(cb), n=np.stack(nn),
                        tm=tm, rmsd=rm, glocon=glocon)
    print("TMscore golden written:", tm[2, 0], rm[2, 0])


def make_reliability_golden():
    """calculate_reliability_score (utils_trX2dy/utils.py:337-372) on the reference's own example decoys.  Bio.PDB is
    absent here, so the phi/psi list PPBuilder would give (every residue with both angles defined, radians, IUPAC
    sign) is computed from the backbone with a plain dihedral; the SCORE is the reference's own ramachandran_score
    function applied to that list (note its degree bounds on radian angles: it counts phi <= 0)."""
    U = load_reference_geometry()
    import glob
    bbs, scores, names = [], [], []
    for pdb in sorted(glob.glob(os.path.join(REF, "example/output/seq/pred_pdb/conf_*.pdb"))):
        res = {}
        for ln in open(pdb):
            if ln.startswith("ATOM") and ln[12:16].strip() in ("N", "CA", "C"):
                res.setdefault(int(ln[22:26]), {})[ln[12:16].strip()] = [float(ln[30:38]), float(ln[38:46]), float(ln[46:54])]
        keys = sorted(res)
        bb = np.array([[res[k]["N"], res[k]["CA"], res[k]["C"]] for k in keys])

        def dih(p0, p1, p2, p3):
            b0, b1, b2 = p0 - p1, p2 - p1, p3 - p2
            b1 = b1 / np.linalg.norm(b1)
            v, w = b0 - np.dot(b0, b1) * b1, b2 - np.dot(b2, b1) * b1
            return float(np.arctan2(np.dot(np.cross(b1, v), w), np.dot(v, w)))
        L = len(bb)
        lst = [(dih(bb[i - 1, 2], bb[i, 0], bb[i, 1], bb[i, 2]), dih(bb[i, 0], bb[i, 1], bb[i, 2], bb[i + 1, 0])) for i in range(1, L - 1)]
        bbs.append(bb); scores.append(U.ramachandran_score(lst)); names.append(os.path.basename(pdb))
    np.savez_compressed(os.path.join(HERE, "example_reliability.npz"), bb=np.array(bbs), score=np.array(scores), names=np.array(names))
    print("example_reliability.npz", dict(zip(names, scores)))


def make_backbone_stats_golden():
    """Ideal backbone geometry is a [ROSETTA-RECALL] item.  The reference's example decoys were
    written by Rosetta after IdealizeMover + a final minimisation, so their bond lengths and angles sit on
    Rosetta's ideal values: keep their means / standard deviations as the fixture the constants of
    include/trx_centroid_model.h are checked against."""
    import glob

    def ang(a, b, c):
        u, v = a - b, c - b
        return np.degrees(np.arccos(np.dot(u, v) / np.linalg.norm(u) / np.linalg.norm(v)))

    def dih(p1, p2, p3, p4):
        b0, b1, b2 = p1 - p2, p3 - p2, p4 - p3
        b1 = b1 / np.linalg.norm(b1)
        v, w = b0 - np.dot(b0, b1) * b1, b2 - np.dot(b2, b1) * b1
        return np.degrees(np.arctan2(np.dot(np.cross(b1, v), w), np.dot(v, w)))
    keys = ("N-CA", "CA-C", "C-N", "C-O", "CA-CB", "N-CA-C", "CA-C-N", "C-N-CA", "CA-C-O", "O-C-N", "abs_omega", "N-C-CA-CB")
    S = {k: [] for k in keys}
    for f in sorted(glob.glob(f"{REF}/example/output/seq/pred_pdb/conf_*.pdb")):
        at = {}
        for ln in open(f):
            if ln.startswith("ATOM") and ln[12:16].strip() in ("N", "CA", "C", "O", "CB"):
                at.setdefault(int(ln[22:26]), {})[ln[12:16].strip()] = np.array([float(ln[30:38]), float(ln[38:46]), float(ln[46:54])])
        r = [at[i] for i in sorted(at)]
        for i, a in enumerate(r):
            S["N-CA"].append(np.linalg.norm(a["CA"] - a["N"])); S["CA-C"].append(np.linalg.norm(a["C"] - a["CA"]))
            S["C-O"].append(np.linalg.norm(a["O"] - a["C"]))
            S["N-CA-C"].append(ang(a["N"], a["CA"], a["C"])); S["CA-C-O"].append(ang(a["CA"], a["C"], a["O"]))
            if "CB" in a:
                S["CA-CB"].append(np.linalg.norm(a["CB"] - a["CA"])); S["N-C-CA-CB"].append(dih(a["N"], a["C"], a["CA"], a["CB"]))
            if i + 1 < len(r):
                b = r[i + 1]
                S["C-N"].append(np.linalg.norm(b["N"] - a["C"])); S["CA-C-N"].append(ang(a["CA"], a["C"], b["N"]))
                S["C-N-CA"].append(ang(a["C"], b["N"], b["CA"])); S["O-C-N"].append(ang(a["O"], a["C"], b["N"]))
                S["abs_omega"].append(abs(dih(a["CA"], a["C"], b["N"], b["CA"])))
    np.savez_compressed(f"{HERE}/example_backbone_stats.npz", keys=np.array(keys), mean=np.array([np.mean(S[k]) for k in keys]),
                        sd=np.array([np.std(S[k]) for k in keys]), n=np.array([len(S[k]) for k in keys]))
    print("backbone statistics written")


if __name__ == "__main__":
    if "--backbone-only" in sys.argv:
        make_backbone_stats_golden()
        sys.exit(0)
    if "--tmscore-only" in sys.argv:
        make_tmscore_golden()
        sys.exit(0)
    if "--variants-only" in sys.argv:
        make_variant_golden()
        sys.exit(0)
    if "--dynamics-only" not in sys.argv:
        main()
        make_variant_golden()
        make_tmscore_golden()
        make_backbone_stats_golden()
    make_dynamics_golden()
