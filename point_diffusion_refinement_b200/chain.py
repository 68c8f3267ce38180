"""Host side of the fused stage kernel (csrc/stage_chain.cu, ``pdr_stage_chain``): ctypes mirrors of the
``PdrChain*`` structs, the shared-memory weight image, the TMEM column plan and the step programs (sweeps) of one
grouped stage -- ``Mlp_plus_t_emb`` over grouped rows + ``AttentionModule`` pooling
(reference: pointnet2_ops/pointnet2_modules.py:57-65,129-174, attention.py:70-96).

A stage with L MLP layers is evaluated in L + 2 sweeps over its grouped rows; sweep d recomputes the chain from the
gathered rows X0 up to depth d with every intermediate kept in tensor memory, and emits either the per-tile statistics the
next GroupNorm needs or (last sweep) the pooled rows:

    sweep 1      [y1 | key] = X0.W1^T                                  -> stats(y1), relu-stats(key)
    sweep 2      a1 = relu(gn(y1)) + e1, k1 = gn(relu(key))
                 y2 = a1.W2^T,  s1 = k1.W1k^T + query row              -> stats(y2), relu-stats(s1)
    sweep d<=L   ... y_d = a_{d-1}.W_d^T                               -> stats(y_d)
    sweep L+1    V = a_L.Wv^T + X0.(Wv.Wres)^T                         -> stats(V)
    sweep L+2    everything + S = gn(relu(s1)).Ws^T                    -> out = sum_k softmax_k(S) relu(gn(V))

``emulate_sweep`` executes a sweep in numpy exactly as the kernel does (same step program, same weight image, same TMEM
columns); tests/test_chain_host.py holds it against a direct evaluation of the stage, so the planner is verified
without a GPU.  The product path never calls it.
"""
import ctypes

import numpy as np
import torch

c_void = ctypes.c_void_p
MAX_STEPS, MAX_MMA, MAX_EPI = 4, 4, 3
XFORM, STATS, POOL = 1, 2, 3
PRO_NONE, PRO_GN_RELU, PRO_RELU_GN = 0, 1, 2
GROUP_COLS = 256          # TMEM columns of one tile group (stage_chain.cu kGroupCols)
MAX_STAT_COLS = 128
MAX_CONST_COLS = 512       # 32-padded columns of all epilogue operations of one sweep (stage_chain.cu kConstCols)
TILE_ROWS = 128


class ChainMma(ctypes.Structure):
    _fields_ = [("d_col", ctypes.c_int), ("n", ctypes.c_int), ("a_tmem", ctypes.c_int), ("a_col", ctypes.c_int),
                ("k", ctypes.c_int), ("w_off", ctypes.c_int), ("w_rows", ctypes.c_int), ("w_row0", ctypes.c_int),
                ("w_k0", ctypes.c_int), ("accumulate", ctypes.c_int)]


class ChainEpi(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("d_col", ctypes.c_int), ("ncols", ctypes.c_int),
                ("bias", c_void), ("rowadd", c_void), ("ld_rowadd", ctypes.c_int),
                ("pro_mode", ctypes.c_int), ("sc", c_void), ("sh", c_void), ("ld_scsh", ctypes.c_int),
                ("emb", c_void), ("ld_emb", ctypes.c_int), ("a_col", ctypes.c_int),
                ("stat_col0", ctypes.c_int), ("stat_skip", ctypes.c_int),
                ("v_col", ctypes.c_int), ("v_bias", c_void), ("v_sc", c_void), ("v_sh", c_void),
                ("v_ld_scsh", ctypes.c_int)]


class ChainStep(ctypes.Structure):
    _fields_ = [("n_mma", ctypes.c_int), ("n_epi", ctypes.c_int), ("release_x0", ctypes.c_int),
                ("mma", ChainMma * MAX_MMA), ("epi", ChainEpi * MAX_EPI)]


class ChainArgs(ctypes.Structure):
    _fields_ = [("table", c_void), ("ld_table", ctypes.c_int), ("k_split", ctypes.c_int),
                ("src_rows", c_void), ("geo", c_void), ("ld_geo", ctypes.c_int), ("k0", ctypes.c_int),
                ("w_image", c_void), ("w_bytes", ctypes.c_int),
                ("batch", ctypes.c_int), ("rows_per_sample", ctypes.c_int), ("group_k", ctypes.c_int),
                ("stats", c_void), ("stats_n", ctypes.c_int), ("stats_relu_mask", ctypes.c_uint * 4),
                ("counts", c_void), ("out", c_void), ("ld_out", ctypes.c_int), ("max_ctas", ctypes.c_int), ("round_out", ctypes.c_int),
                ("n_steps", ctypes.c_int), ("steps", ChainStep * MAX_STEPS)]


def p32(c):
    return (c + 31) // 32 * 32


def r4(c):
    return (c + 3) // 4 * 4


def tf32_round(w):
    """Round fp32 to the nearest TF32 value (ties away from zero): the tensor core truncates what it is given."""
    bits = w.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


class WeightImage:
    """Byte image of every weight matrix of a stage in the layout stage_chain.cu copies into shared memory: per matrix
    (rows padded to 32, K padded to 32) chunk kc = columns [32 kc, 32 kc + 32) is `rows` rows of 128 bytes, the 16-byte
    piece p of row r stored at piece p ^ (r & 7) -- the K-major SWIZZLE_128B operand layout of tcgen05.mma."""

    def __init__(self):
        self.parts, self.bytes, self.index = [], 0, {}

    def add(self, name, w, round_tf32=True):
        """w: torch (N, K) fp32.  Returns (byte offset, padded rows)."""
        w = w.detach().float().cpu()
        if round_tf32:
            w = tf32_round(w)
        n, k = w.shape
        rows, kp = p32(n), p32(k)
        full = torch.zeros(rows, kp)
        full[:n, :k] = w
        a = full.numpy().reshape(rows, kp // 32, 8, 4)                  # (row, chunk, piece, 4 floats)
        img = np.zeros((kp // 32, rows, 8, 4), dtype=np.float32)
        r = np.arange(rows)
        for p in range(8):
            img[:, r, p ^ (r & 7), :] = a[:, :, p, :].transpose(1, 0, 2)
        off = self.bytes
        self.parts.append(img.reshape(-1))
        self.bytes += img.size * 4
        assert off % 1024 == 0 and self.bytes % 1024 == 0
        self.index[name] = (off, rows, kp)
        return off, rows

    def tensor(self, device):
        return torch.from_numpy(np.concatenate(self.parts)).to(device)

    @staticmethod
    def decode(image, off, rows, kp):
        """(rows, kp) matrix back out of an image (numpy float32 1-D) -- used by the emulator."""
        img = image[off // 4: off // 4 + rows * kp].reshape(kp // 32, rows, 8, 4)
        out = np.zeros((rows, kp // 32, 8, 4), dtype=np.float32)
        r = np.arange(rows)
        for p in range(8):
            out[:, :, p, :] = img[:, r, p ^ (r & 7), :].transpose(1, 0, 2)
        return out.reshape(rows, kp)


class TmemPlan:
    """First-fit allocation of TMEM column ranges (units of 32 columns, GROUP_COLS per tile group) to tensors with
    lifetimes on the step clock: MMA phase of step s = 2 s, epilogue phase = 2 s + 1 (both inclusive)."""

    def __init__(self):
        self.items = []       # (col0, ncols, t0, t1, name)

    def alloc(self, name, ncols, t0, t1):
        ncols = p32(ncols)
        for col0 in range(0, GROUP_COLS - ncols + 1, 32):
            ok = True
            for (c, n, a, b, _) in self.items:
                if col0 < c + n and c < col0 + ncols and t0 <= b and a <= t1:
                    ok = False
                    break
            if ok:
                self.items.append((col0, ncols, t0, t1, name))
                return col0
        raise MemoryError("TMEM plan: no room for %s (%d columns, steps %d..%d)" % (name, ncols, t0, t1))


class StageSpec:
    """Static description of one grouped stage (weights already mapped onto the X0 column layout).

    k0 / k_split: columns of the gathered operand / of its table part;  mlp: [(W, b)] for layers 1..L, layer 1 of
    shape (c1, k0);  key, w1k, ws: (W, b) of grouped_feat_conv, the key part of weight_conv[2], weight_conv[5];
    wv: (Wv (co, cL), Wvt (co, k0), bv) = feat_out_conv with the residual convolution folded into the X0 tail."""

    def __init__(self, k0, k_split, mlp, key, w1k, ws, wv):
        self.k0, self.k_split, self.mlp, self.key, self.w1k, self.ws, self.wv = k0, k_split, mlp, key, w1k, ws, wv
        self.L = len(mlp)
        assert self.L >= 2, "Mlp_plus_t_emb has at least first_mlp and second_mlp"
        self.c = [w.shape[0] for w, _ in mlp]                  # c[l-1] = width of layer l
        self.ck, self.ci, self.co = key[0].shape[0], w1k[0].shape[0], ws[0].shape[0]
        assert mlp[0][0].shape[1] == k0 and key[0].shape[1] == k0 and w1k[0].shape[1] == self.ck
        assert ws[0].shape[1] == self.ci and wv[0].shape == (self.co, self.c[-1]) and wv[1].shape == (self.co, k0)
        for l in range(1, self.L):
            assert mlp[l][0].shape[1] == self.c[l - 1]


def _pad_bias(b, n, device):
    out = torch.zeros(p32(n), device=device)
    if b is not None:
        out[:b.numel()] = b.detach().float().to(device)
    return out


class StagePlan:
    """Weight image, padded biases and eligibility of a stage for pdr_stage_chain."""

    def __init__(self, spec, device):
        self.spec, self.dev = spec, device
        s = spec
        img = WeightImage()
        c1p, ckp = p32(s.c[0]), p32(s.ck)
        k0p = p32(s.k0)
        w1cat = torch.zeros(c1p + ckp, s.k0)
        w1cat[:s.c[0]] = s.mlp[0][0].detach().float().cpu()
        w1cat[c1p:c1p + s.ck] = s.key[0].detach().float().cpu()
        self.w1 = img.add("w1cat", w1cat)
        self.wl = [None] + [img.add("w%d" % (l + 1), s.mlp[l][0]) for l in range(1, s.L)]
        self.w1k = img.add("w1k", s.w1k[0])
        self.ws = img.add("ws", s.ws[0])
        self.wv = img.add("wv", s.wv[0])
        self.wvt = img.add("wvt", s.wv[1])
        self.image = img
        self.image_t = img.tensor(device)
        self.k0p = k0p
        self.k0m = (s.k0 + 7) // 8 * 8          # K of the MMAs that read X0 (its columns >= k0 are zero-filled)
        self.b = {"y1": _pad_bias(s.mlp[0][1], s.c[0], device), "key": _pad_bias(s.key[1], s.ck, device),
                  "s1": _pad_bias(s.w1k[1], s.ci, device), "S": _pad_bias(s.ws[1], s.co, device),
                  "V": _pad_bias(s.wv[2], s.co, device)}
        for l in range(1, s.L):
            self.b["y%d" % (l + 1)] = _pad_bias(s.mlp[l][1], s.c[l], device)

    def smem_bytes(self, slots=2, wpg=4):
        """Dynamic shared memory of a launch with `slots` X0 tiles in the ring and `wpg` epilogue warps per tile group
        (stage_chain.cu: transposition tiles + weight image + ring, + alignment slack)."""
        return 1024 + 2 * wpg * 32 * 36 * 4 + self.image.bytes + slots * (self.k0p // 32) * TILE_ROWS * 128

    def fits(self):
        """Shared memory (weights + at least two X0 tiles with 4 epilogue warps per group), statistics columns, folded-constant
        columns and MMA widths within the kernel's limits."""
        static = 25600 + 512                   # stage_chain.cu: barriers, column partials, folded constants, decoded program
        s = self.spec
        const_cols = 2 * p32(s.co) + sum(p32(c) for c in s.c) + p32(s.ck) + p32(s.ci)       # the last sweep holds them all
        return (self.smem_bytes(2, 4) <= 227 * 1024 - static and p32(s.c[0]) + p32(s.ck) <= 256
                and const_cols <= MAX_CONST_COLS
                and all(self.sweep_stats_n(d) <= MAX_STAT_COLS for d in range(1, s.L + 2)))

    # ---- step programs --------------------------------------------------------------------------------
    def n_sweeps(self):
        return self.spec.L + 2

    def sweep_stat_columns(self, d):
        """[(name, stat_col0, ncols, relu)] written by sweep d (1-based); empty for the last sweep."""
        s = self.spec
        if d == 1:
            return [("y1", 0, s.c[0], False), ("key", p32(s.c[0]), s.ck, True)]
        if d == 2:
            return [("y2", 0, s.c[1], False), ("s1", p32(s.c[1]), s.ci, True)]
        if d <= s.L:
            return [("y%d" % d, 0, s.c[d - 1], False)]
        if d == s.L + 1:
            return [("V", 0, s.co, False)]
        return []

    def build_sweep(self, d, rt):
        """ChainArgs of sweep d.  rt: runtime bindings --
             table (ptr, ld), src_rows ptr, geo (ptr, ld), batch, rows_per_sample, group_k,
             stats: ptr of this sweep's (tiles, stats_n, 4) buffer (sweeps 1..L+1),
             gn["y1"].."y<L>", gn["key"], gn["s1"], gn["V"]: (sc ptr, sh ptr, ld) as far as sweep d needs them,
             emb[l] (l = 1..L): (ptr, ld) or None -- added after relu(gn(y_l)),
             rowadd: (ptr, ld) of the query rows (sweeps >= 2), counts ptr|None, out (ptr, ld) (last sweep).
        Returns (args, steps) where steps is a python description [(mmas, epis, release)] used by the emulator."""
        s = self.spec
        L = s.L
        last = d == L + 2
        c1p, ckp = p32(s.c[0]), p32(s.ck)
        need_scores = d == 2 or last            # key -> s1 (-> S)
        need_key_stats = d == 1
        # depth of the MLP branch evaluated in this sweep: layers 1..m, then V if d >= L + 1
        m = min(d, L)
        want_v = d >= L + 1
        # ---- step list (python form): each step = dict(mma=[...], epi=[...]) --------------------------------
        steps = [dict(mma=[], epi=[]) for _ in range(MAX_STEPS)]
        plan = TmemPlan()
        n_steps = (L + 1) if want_v else m           # V sits at step L (0-based); y_m at step m - 1
        if need_scores:
            n_steps = max(n_steps, 3 if last else 2)
        assert n_steps <= MAX_STEPS
        v_step = L                                   # step whose MMA completes V
        # Where the X0 tail of V is contracted: with the first step (the X0 tile is released at once) if TMEM has room
        # for an accumulator that lives through the whole tile, otherwise with V's own step (the tile is held)
        y1_n = c1p + (ckp if (need_scores or need_key_stats) else 0)
        cols = {}

        def stat_of(name):
            for nm, c0, nc, relu in self.sweep_stat_columns(d):
                if nm == name:
                    return c0, relu
            return None

        def add_layer_out(name, step, width, src_cols, bias, stats_wanted, xform=None):
            """Epilogue of one accumulator: STATS if this sweep emits its statistics, else XFORM (if consumed)."""
            st = stat_of(name)
            if st is not None and stats_wanted:
                steps[step]["epi"].append(dict(kind=STATS, d_col=src_cols, ncols=width, bias=bias, stat_col0=st[0],
                                               stat_skip=(1 if st[1] else 2), rowadd=(name == "s1")))
            elif xform is not None:
                steps[step]["epi"].append(dict(kind=XFORM, d_col=src_cols, ncols=width, bias=bias, rowadd=(name == "s1"),
                                               **xform))

        early_tail = False
        if want_v:
            try:
                trial = TmemPlan()
                trial.alloc("V", s.co, 0, 2 * v_step + 1)
                self._alloc_rest(trial, d, m, need_scores, last, y1_n, v_late=False)
                early_tail = True
            except MemoryError:
                early_tail = False
        if want_v and early_tail:
            cols["V"] = plan.alloc("V", s.co, 0, 2 * v_step + 1)
        cols.update(self._alloc_rest(plan, d, m, need_scores, last, y1_n, v_late=want_v and not early_tail))
        # ---- step 0: X0 -> [y1 | key] (+ V tail) ------------------------------------------------------------
        steps[0]["mma"].append(dict(d_col=cols["Y1"], n=y1_n, a_tmem=0, a_col=0, k=self.k0m, w=self.w1, w_row0=0, w_k0=0,
                                    accumulate=0))
        if want_v and early_tail:
            steps[0]["mma"].append(dict(d_col=cols["V"], n=p32(s.co), a_tmem=0, a_col=0, k=self.k0m, w=self.wvt, w_row0=0,
                                        w_k0=0, accumulate=0))
        release_step = 0 if (not want_v or early_tail) else v_step
        # y1
        if m >= 2 or want_v:
            xf = dict(pro_mode=PRO_GN_RELU, gn="y1", emb=1, a_col=cols["A1"])
        else:
            xf = None
        add_layer_out("y1", 0, s.c[0], cols["Y1"], "y1", d == 1, xf)
        if need_key_stats:
            add_layer_out("key", 0, s.ck, cols["Y1"] + c1p, "key", True)
        elif need_scores:
            add_layer_out("key", 0, s.ck, cols["Y1"] + c1p, "key", False,
                          dict(pro_mode=PRO_RELU_GN, gn="key", emb=None, a_col=cols["K1"]))
        # ---- MLP layers 2..m -----------------------------------------------------------------------------
        for l in range(2, m + 1):
            st = l - 1
            steps[st]["mma"].append(dict(d_col=cols["Y%d" % l], n=p32(s.c[l - 1]), a_tmem=1, a_col=cols["A%d" % (l - 1)],
                                         k=p32(s.c[l - 2]), w=self.wl[l - 1], w_row0=0, w_k0=0, accumulate=0))
            xf = dict(pro_mode=PRO_GN_RELU, gn="y%d" % l, emb=l, a_col=cols.get("A%d" % l)) if (l < m or want_v) else None
            add_layer_out("y%d" % l, st, s.c[l - 1], cols["Y%d" % l], "y%d" % l, d == l, xf)
        # ---- score branch --------------------------------------------------------------------------------
        if need_scores:
            steps[1]["mma"].append(dict(d_col=cols["S1"], n=p32(s.ci), a_tmem=1, a_col=cols["K1"], k=ckp, w=self.w1k,
                                        w_row0=0, w_k0=0, accumulate=0))
            xf = dict(pro_mode=PRO_RELU_GN, gn="s1", emb=None, a_col=cols.get("S1n")) if last else None
            add_layer_out("s1", 1, s.ci, cols["S1"], "s1", d == 2, xf)
            if last:
                steps[2]["mma"].append(dict(d_col=cols["S"], n=p32(s.co), a_tmem=1, a_col=cols["S1n"], k=p32(s.ci),
                                            w=self.ws, w_row0=0, w_k0=0, accumulate=0))
        # ---- values --------------------------------------------------------------------------------------
        if want_v:
            if not early_tail:
                steps[v_step]["mma"].append(dict(d_col=cols["V"], n=p32(s.co), a_tmem=0, a_col=0, k=self.k0m, w=self.wvt,
                                                 w_row0=0, w_k0=0, accumulate=0))
            steps[v_step]["mma"].append(dict(d_col=cols["V"], n=p32(s.co), a_tmem=1, a_col=cols["A%d" % L], k=p32(s.c[-1]),
                                             w=self.wv, w_row0=0, w_k0=0, accumulate=1))
            if last:
                steps[v_step]["epi"].append(dict(kind=POOL, d_col=cols["S"], ncols=s.co, bias="S", v_col=cols["V"]))
            else:
                add_layer_out("V", v_step, s.co, cols["V"], "V", True)
        steps = steps[:n_steps]
        for i, st in enumerate(steps):
            st["release"] = int(i == release_step)
            assert 1 <= len(st["mma"]) <= MAX_MMA and 1 <= len(st["epi"]) <= MAX_EPI, (d, i, st)
        return self._encode(d, steps, rt), steps

    def _alloc_rest(self, plan, d, m, need_scores, last, y1_n, v_late):
        """Column ranges of every tensor of sweep d except an early V accumulator.  Lifetimes on the step clock."""
        s = self.spec
        L = s.L
        cols = {}
        want_v = d >= L + 1
        cols["Y1"] = plan.alloc("Y1", y1_n, 0, 1)
        if m >= 2 or want_v:
            cols["A1"] = plan.alloc("A1", s.c[0], 1, 2)
        if need_scores:
            cols["K1"] = plan.alloc("K1", s.ck, 1, 2)
        for l in range(2, m + 1):
            st = l - 1
            cols["Y%d" % l] = plan.alloc("Y%d" % l, s.c[l - 1], 2 * st, 2 * st + 1)
            if l < m or want_v:
                cols["A%d" % l] = plan.alloc("A%d" % l, s.c[l - 1], 2 * st + 1, 2 * st + 2)
        if need_scores:
            cols["S1"] = plan.alloc("S1", s.ci, 2, 3)
            if last:
                cols["S1n"] = plan.alloc("S1n", s.ci, 3, 4)
                cols["S"] = plan.alloc("S", s.co, 4, 2 * L + 1)
        if v_late:
            cols["V"] = plan.alloc("V", s.co, 2 * L, 2 * L + 1)
        return cols

    def _encode(self, d, steps, rt):
        s = self.spec
        a = ChainArgs()
        a.table, a.ld_table, a.k_split = rt["table"][0], rt["table"][1], s.k_split
        a.src_rows, a.geo, a.ld_geo, a.k0 = rt["src_rows"], rt["geo"][0], rt["geo"][1], s.k0
        a.w_image, a.w_bytes = self.image_t.data_ptr(), self.image.bytes
        a.batch, a.rows_per_sample, a.group_k = rt["batch"], rt["rows_per_sample"], rt["group_k"]
        stat_cols = self.sweep_stat_columns(d)
        if stat_cols:
            a.stats = rt["stats"]
            a.stats_n = self.sweep_stats_n(d)
            for _, c0, nc, relu in stat_cols:
                if relu:
                    for c in range(c0, c0 + p32(nc)):
                        a.stats_relu_mask[c >> 5] |= (1 << (c & 31))
        else:
            a.counts = rt.get("counts")
            a.out, a.ld_out = rt["out"]
        a.max_ctas = rt.get("max_ctas", 0)
        a.round_out = rt.get("round_out", 0)
        a.n_steps = len(steps)
        for i, st in enumerate(steps):
            cs = a.steps[i]
            cs.n_mma, cs.n_epi, cs.release_x0 = len(st["mma"]), len(st["epi"]), st["release"]
            for j, m in enumerate(st["mma"]):
                o = cs.mma[j]
                o.d_col, o.n, o.a_tmem, o.a_col, o.k = m["d_col"], m["n"], m["a_tmem"], m["a_col"], m["k"]
                o.w_off, o.w_rows, o.w_row0, o.w_k0, o.accumulate = m["w"][0], m["w"][1], m["w_row0"], m["w_k0"], m["accumulate"]
            for j, e in enumerate(st["epi"]):
                o = cs.epi[j]
                o.kind, o.d_col, o.ncols = e["kind"], e["d_col"], e["ncols"]
                o.bias = self.b[e["bias"]].data_ptr()
                if e.get("rowadd"):
                    o.rowadd, o.ld_rowadd = rt["rowadd"]
                if e["kind"] == XFORM:
                    sc, sh, ld = rt["gn"][e["gn"]]
                    o.pro_mode, o.sc, o.sh, o.ld_scsh, o.a_col = e["pro_mode"], sc, sh, ld, e["a_col"]
                    emb = rt.get("emb", {}).get(e["emb"]) if e["emb"] is not None else None
                    if emb is not None:
                        o.emb, o.ld_emb = emb
                elif e["kind"] == STATS:
                    o.stat_col0, o.stat_skip = e["stat_col0"], e["stat_skip"]
                else:
                    sc, sh, ld = rt["gn"]["V"]
                    o.v_col, o.v_bias, o.v_sc, o.v_sh, o.v_ld_scsh = e["v_col"], self.b["V"].data_ptr(), sc, sh, ld
        return a

    def sweep_stats_n(self, d):
        cols = self.sweep_stat_columns(d)
        return max(c0 + p32(nc) for _, c0, nc, _ in cols) if cols else 0


# ----------------------------------------------------------------------------------------------------------
# numpy execution of a sweep, step for step as stage_chain.cu does it (tests only)
# ----------------------------------------------------------------------------------------------------------
def emulate_sweep(plan, d, steps, host):
    """host: numpy arrays -- X0 (M, k0), batch, rows_per_sample, group_k, gn[name] = (sc (B, ld), sh), emb[l] (B, c) or
    None, rowadd (points, ci), counts (points) or None.  Returns stats (tiles, stats_n, 4) or out (points, co)."""
    s = plan.spec
    image = plan.image_t.cpu().numpy()
    X0 = host["X0"].astype(np.float64)
    M = X0.shape[0]
    K = host["group_k"]
    tiles = M // TILE_ROWS
    tps = host["rows_per_sample"] // TILE_ROWS
    stat_n = plan.sweep_stats_n(d)
    stats = np.zeros((tiles, max(stat_n, 1), 4))
    out = np.zeros((M // K, s.co))
    bias = {k: v.cpu().numpy().astype(np.float64) for k, v in plan.b.items()}
    for t in range(tiles):
        b = t // tps
        rows = slice(t * TILE_ROWS, (t + 1) * TILE_ROWS)
        x0 = np.zeros((TILE_ROWS, plan.k0p))
        x0[:, :s.k0] = X0[rows]
        tmem = np.full((TILE_ROWS, GROUP_COLS), np.nan)
        points = np.arange(t * TILE_ROWS, (t + 1) * TILE_ROWS) // K
        for st in steps:
            for m in st["mma"]:
                off, wrows = m["w"]
                kp = [v[2] for v in plan.image.index.values() if v[0] == off][0]
                W = WeightImage.decode(image, off, wrows, kp).astype(np.float64)
                Wsub = W[m["w_row0"]:m["w_row0"] + m["n"], m["w_k0"]:m["w_k0"] + m["k"]]
                A = tmem[:, m["a_col"]:m["a_col"] + m["k"]] if m["a_tmem"] else x0[:, m["a_col"]:m["a_col"] + m["k"]]
                assert not np.isnan(A).any(), ("A operand reads unwritten TMEM", d, m)
                acc = A @ Wsub.T
                dst = slice(m["d_col"], m["d_col"] + m["n"])
                if m["accumulate"]:
                    assert not np.isnan(tmem[:, dst]).any()
                    tmem[:, dst] += acc
                else:
                    tmem[:, dst] = acc
            for e in st["epi"]:          # in place and in order, as each epilogue warp does for its rows
                nc, ncp = e["ncols"], p32(e["ncols"])
                y = tmem[:, e["d_col"]:e["d_col"] + ncp].copy()
                assert not np.isnan(y).any(), ("epilogue reads unwritten TMEM", d, e)
                y += bias[e["bias"]][:ncp]
                if e.get("rowadd"):
                    y[:, :nc] += host["rowadd"][points][:, :nc]
                if e["kind"] == XFORM:
                    sc, sh = host["gn"][e["gn"]]
                    scv = np.zeros(ncp); shv = np.zeros(ncp); ev = np.zeros(ncp)
                    scv[:nc], shv[:nc] = sc[b, :nc], sh[b, :nc]
                    emb = host.get("emb", {}).get(e["emb"]) if e["emb"] is not None else None
                    if emb is not None:
                        ev[:nc] = emb[b, :nc]
                    if e["pro_mode"] == PRO_GN_RELU:
                        tt = np.maximum(y * scv + shv, 0) + ev
                    else:
                        tt = np.maximum(y, 0) * scv + shv + ev
                    tt[:, nc:] = 0
                    tmem[:, e["a_col"]:e["a_col"] + ncp] = tt
                elif e["kind"] == STATS:
                    c0 = e["stat_col0"]
                    if e["stat_skip"] & 1:
                        p = np.maximum(y, 0)
                        stats[t, c0:c0 + ncp, 2] = p.sum(0); stats[t, c0:c0 + ncp, 3] = (p * p).sum(0)
                    else:
                        stats[t, c0:c0 + ncp, 0] = y.sum(0); stats[t, c0:c0 + ncp, 1] = (y * y).sum(0)
                else:
                    sc, sh = host["gn"]["V"]
                    v = tmem[:, e["v_col"]:e["v_col"] + ncp] + bias["V"][:ncp]
                    vv = np.maximum(v[:, :nc] * sc[b, :nc] + sh[b, :nc], 0)
                    sco = y[:, :nc].reshape(TILE_ROWS // K, K, nc)
                    vv = vv.reshape(TILE_ROWS // K, K, nc)
                    pts = points[::K]
                    for i, pt in enumerate(pts):
                        cnt = K if host.get("counts") is None else max(int(host["counts"][pt]), 1)
                        sk = np.where(np.arange(K)[:, None] < cnt, sco[i], -1e9)
                        ex = np.exp(sk - sk.max(0))
                        out[pt] = (ex * vv[i]).sum(0) / ex.sum(0)
    return stats if stat_n else out
