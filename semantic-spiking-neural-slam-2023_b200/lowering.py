"""Lower a built network into the flat device plan executed by the sm_100a kernels.

Host-side counterpart of SURVEY.md §7 step 4.  nengo's builder would emit thousands of
tiny operators (Reset / Copy / DotInc / ElementwiseInc / SimProcess ...) for the
SSP-SLAM graphs (App. A.6, App. B).  Here the *vector-level glue* is collapsed
algebraically instead of being executed:

* every value that exists at the start of a step or is produced by a heavy op is a
  **source column**: the constant 1, input-table nodes, Lowpass filter states (old
  value = one-step delay), decoded outputs of ensemble connections, outputs of the
  device function nodes (grid clean-up, gated correction);
* every value a heavy op consumes is a **sink row**: ensemble inputs, direct neuron
  currents, filter inputs, function-node inputs, learning-rule inputs, probes.  Each
  sink row is a sparse linear combination of source columns, obtained by composing
  pass-through nodes, slices and transforms with ``scipy.sparse`` (so the 112x55
  circular-convolution DFT matrices, ``to_Fourier`` / ``to_SSP`` etc. are folded into
  CSR rows that the kernels evaluate in their prologues);
* sinks are levelled by data dependency inside a step (level 0 needs only tables and
  old filter states; OVC -> circular convolution is level 1, ...), giving the kernel
  launch order.

The per-trial arena uses ``[row][trial]`` layout everywhere (trial is the coalesced axis).
"""
from __future__ import annotations

import dataclasses
import os
import numpy as np
import scipy.sparse as sp

from . import compat
from . import nodeops
from .builder import BuiltModel

SMALL_MAX_DIMS = 4
SMALL_MAX_OUT = 8
CSR_PAD = 8                 # CSR rows are padded to a multiple of this with (row 0, coefficient 0)
DEC_TILE = 8                # decoder output rows are padded to a multiple of this
N_SM = 148                  # B200
MAX_DEC_CHUNKS = 32
DEC_SMEM_BYTES = 96 * 1024  # shared-memory budget of a k_decode chunk (weights + activities)
SMEM_PER_SM = 227 * 1024
PES_CTAS_PER_SM = 4         # k_pes: 116 registers x 128 threads
DEC_CTAS_PER_SM = 5         # k_decode: 96 registers x 128 threads (before its shared-memory limit)
DEC_TC_WIDTH = 64           # k_decode_tc: padded output width / neurons per stage of the tensor-core decoder
DEC_TC_STAGE = 64


FUSE_MAX_NNZ = 400_000   # largest composed level-0 program that is still fused into the end-of-step launch
TABN_BASE = 1 << 28      # marker of 'table row of the next step' in the parity-0 CSR column list (resolved by _entries)


def _best_chunks(n, units, slots_of, k_min, k_max, fixed=8, per_chunk=0.0):
    """Split a neuron range into k chunks so that the launch (units * k CTAs) wastes the least time on
    partial waves: minimise ceil(CTAs / resident slots) * neurons per chunk (+ the split-K reduce, which
    grows with the number of partial sums)."""
    best, best_cost = k_min, None
    for k in range(k_min, k_max + 1):
        per = -(-n // k)
        waves = -(-(units * k) // max(1, slots_of(per)))
        cost = waves * (per + fixed) + per_chunk * k   # + a fixed per-CTA cost (prologue, split-K epilogue)
        if best_cost is None or cost < best_cost:
            best, best_cost = k, cost
    return best

NT_LIF, NT_LIFRATE, NT_RELU = 0, 1, 2


def _ens_to_neurons(c):
    return compat.is_ensemble(c.pre_obj) and compat.is_neurons(c.post_obj)


def _idx(key, size):
    return np.atleast_1d(np.arange(size)[key])


def _place(rows, n_rows, mat):
    """Scatter the rows of ``mat`` to positions ``rows`` of an (n_rows x ncols) matrix."""
    if len(rows) == n_rows and mat.shape[0] == n_rows and np.array_equal(rows, np.arange(n_rows)):
        return mat                      # the whole input, in order (most connections): nothing to scatter
    sel = sp.csr_matrix((np.ones(len(rows)), (rows, np.arange(len(rows)))), shape=(n_rows, len(rows)))
    return sel @ mat


def _apply_transform(transform, mat):
    if transform is None:
        return mat
    t = np.asarray(transform, dtype=np.float64)
    if t.ndim == 0:
        return mat * float(t)
    if t.ndim == 1:
        return sp.diags(t) @ mat
    return sp.csr_matrix(t) @ mat


@dataclasses.dataclass
class ProbeInfo:
    probe: object
    kind: str                 # "rows" | "weights" | "scaled_encoders"
    row0: int = 0             # first row in the probe buffer
    size: int = 0
    period: int = 1
    conn: object = None
    ens: object = None


class DevicePlan:
    """Numpy arrays + scalars handed to the C-ABI (``ssb_set_array`` / ``ssb_set_scalar``)."""

    def __init__(self):
        self.arrays: dict[str, np.ndarray] = {}
        self.scalars: dict[str, float] = {}
        self.tables: list = []          # [(node, col0, size)]
        self.probes: list[ProbeInfo] = []
        self.ens_state: dict = {}       # ens -> (state_row0, n)
        self.ens_act: dict = {}         # ens -> act_row0 (big ensembles only)
        self.learned_enc: dict = {}     # ens -> (row0, n, dims)   rows are n*dims + k
        self.learned_dec: dict = {}     # conn -> (row0, size_out, n) rows are j*n + i
        self.static_dec: dict = {}      # conn -> (w_off, size_out, jpad, n) for big decoders
        self.pt_enc: dict = {}          # per-trial plans: static wide ens -> (lenc row0, n, dims)
        self.pt_dec: dict = {}          # per-trial plans: static wide decoder -> (ldec row0, size_out, n)
        self.filters: dict = {}         # conn/probe -> (filter_row0, size)
        self.launches: list = []        # human-readable launch order
        self.stats: dict = {}
        self.stats_fusion: dict = {}    # nnz of level 0 / filter updates / their composition (step fusion), if considered


class _Lowerer:
    def __init__(self, network, model: BuiltModel, n_trials=1, per_trial=False):
        self.net, self.model, self.dt = network, model, model.dt
        # every trial has its own network seed: seed-dependent static weights go to per-trial arenas (see lower())
        self.per_trial = bool(per_trial)
        self.n_groups = max(1, -(-int(n_trials) // 32))
        self.nodes = network.all_nodes
        self.ensembles = network.all_ensembles
        self.conns = network.all_connections
        self.probes = network.all_probes
        self.owners = [network] + network.all_networks
        self.incoming: dict = {}
        for c in self.conns:
            self.incoming.setdefault(c.post_obj, []).append(c)
        self.plan = DevicePlan()

    # ------------------------------------------------------------------ classification
    def classify_nodes(self):
        self.node_kind, self.node_op = {}, {}
        for node in self.nodes:
            out = node.output
            if out is None:
                kind = "pass"
            elif isinstance(out, np.ndarray):
                kind = "const"
            elif node.size_in == 0:
                kind = "table"
            else:
                op = nodeops.recognize(node, self.owners)
                if op is None:
                    raise NotImplementedError(
                        f"{node!r}: Python callable with inputs is not a recognised device op "
                        "(identity / grid clean-up / gated correction); the B200 backend has no host-callback path")
                self.node_op[node] = op
                kind = {"identity": "pass", "cleanup": "fn", "gate": "fn"}[op.kind]
            self.node_kind[node] = kind

    # ------------------------------------------------------------------ ensemble classes
    def classify_ensembles(self):
        """Narrow ensembles (VCOs, product squares) are fused encode->neuron->decode items; wide ones
        write activities and are decoded by separate launches whose neuron range is split into chunks
        (split-K: partial sums are parked in a scratch arena and added up by the last CTA to arrive)."""
        self.ens_dec_conns = {e: [] for e in self.ensembles}
        for conn in self.conns:
            if compat.is_ensemble(conn.pre_obj):
                self.ens_dec_conns[conn.pre_obj].append(conn)
        for probe in self.probes:
            if compat.is_ensemble(probe.obj) and probe.attr == "decoded_output":
                self.ens_dec_conns[probe.obj].append(probe)
        voja_posts = {c.post_obj for c in self.conns
                      if c.learning_rule is not None and compat.rule_kind(c.learning_rule.learning_rule_type) == "voja"}
        pes_conns = {c for c in self.conns
                     if c.learning_rule is not None and compat.rule_kind(c.learning_rule.learning_rule_type) == "pes"}
        jn_posts = {c.post_obj.ensemble for c in self.conns if compat.is_neurons(c.post_obj)}
        # ensembles whose neuron outputs are probed keep their activities in the act arena (wide path)
        neuron_probed = {p.obj.ensemble for p in self.probes if compat.is_neurons(p.obj)}
        self.is_small, self.dec_chunks = {}, {}
        for ens in self.ensembles:
            outs = self.ens_dec_conns[ens]
            nout = sum(self._out_size(c) for c in outs)
            has_pes = any(c in pes_conns for c in outs)
            small = (ens.dimensions <= SMALL_MAX_DIMS and nout <= SMALL_MAX_OUT and ens not in voja_posts
                     and ens not in jn_posts and not has_pes and ens not in neuron_probed)
            self.is_small[ens] = small
        n_static = sum(1 for e in self.ensembles if not self.is_small[e]
                       for c in self.ens_dec_conns[e] if c not in pes_conns)
        max_jpad = max([-(-self._out_size(c) // DEC_TILE) * DEC_TILE for e in self.ensembles if not self.is_small[e]
                        for c in self.ens_dec_conns[e] if c not in pes_conns] + [0])
        for ens in self.ensembles:
            for c in self.ens_dec_conns[ens]:
                if self.is_small[ens]:
                    self.dec_chunks[c] = 1
                elif c in pes_conns:   # k_pes: CTA = (8-row tile, trial group, neuron chunk)
                    jtiles = -(-self._out_size(c) // DEC_TILE)
                    k_max = int(max(1, min(ens.n_neurons // 32, MAX_DEC_CHUNKS)))
                    self.dec_chunks[c] = _best_chunks(ens.n_neurons, jtiles * self.n_groups,
                                                      lambda per: N_SM * PES_CTAS_PER_SM, 1, k_max)
                    if os.environ.get("SSB_PES_CHUNKS"):          # tuning knob (scripts/dev_perf.py sweeps)
                        self.dec_chunks[c] = int(max(1, min(int(os.environ["SSB_PES_CHUNKS"]), k_max)))
                elif self.per_trial:   # per-trial static decoders: k_decode_pt walks the whole ensemble per trial
                    self.dec_chunks[c] = 1
                else:                  # static decoders
                    jpad = -(-self._out_size(c) // DEC_TILE) * DEC_TILE
                    quads = -(-self.n_groups // 4)
                    if os.environ.get("SSB_DECODE") != "ffma":
                        # k_decode_tc: CTA = (decoder, tile of 128 output columns, 128 trials, K chunk of 64- (or 32-) neuron
                        # stages), one CTA per SM; decoders wider than 128 columns (d = 649) take several column tiles
                        n_stages = -(-ens.n_neurons // (DEC_TC_STAGE if max_jpad <= DEC_TC_WIDTH else DEC_TC_STAGE // 2))
                        n_nt = 1 if max_jpad <= DEC_TC_WIDTH else -(-max_jpad // (2 * DEC_TC_WIDTH))
                        self.dec_chunks[c] = int(max(1, min(n_stages, N_SM // max(1, n_static * quads * n_nt))))
                        continue
                    # k_decode (FFMA): CTA = (decoder, quad of trial groups, neuron chunk), all outputs at once
                    per_max = max(1, DEC_SMEM_BYTES // (jpad * 4 + 4 * 128))   # weight tile + 4 activity tiles of a chunk
                    need = -(-ens.n_neurons // per_max)
                    k_max = int(max(need, min(max(1, ens.n_neurons // 32), MAX_DEC_CHUNKS)))

                    def slots(per, jpad=jpad):      # CTA = 4 trial groups: shared weight tile + 4 activity tiles
                        smem = per * (jpad * 4 + 4 * 128) + 1024
                        return N_SM * max(1, min(DEC_CTAS_PER_SM, SMEM_PER_SM // smem))
                    quads = -(-self.n_groups // 4)
                    self.dec_chunks[c] = _best_chunks(ens.n_neurons, max(1, n_static) * quads, slots, need, k_max, fixed=12,
                                                      per_chunk=3.0)

    @staticmethod
    def _out_size(c):
        if compat.is_connection(c):
            # ensemble -> neurons (slam_loihi.py:266): decode size_mid values, the (n x size_mid) transform is applied
            # as direct neuron currents after the synapse (linear, so equal to filtering the n folded currents)
            return c.size_mid if _ens_to_neurons(c) else c.size_out
        return c.size_in

    def _dec_expr(self, c):
        """Decoded value of connection/probe ``c`` (split-K partial sums are reduced inside the launch)."""
        return self._eye(self.dec_col[c], self._out_size(c))

    # ------------------------------------------------------------------ source columns
    def enumerate_sources(self):
        self.ncol = 1  # column 0 = constant one
        self.col_kind = ["const"]
        self.col_owner = [None]

        def alloc(kind, owner, size):
            c0 = self.ncol
            self.ncol += size
            self.col_kind += [kind] * size
            self.col_owner += [owner] * size
            return c0

        self.tab_col, self.filt_col, self.dec_col, self.fn_col = {}, {}, {}, {}
        for node in self.nodes:
            if self.node_kind[node] == "table":
                self.tab_col[node] = alloc("tab", node, node.size_out)
        for conn in self.conns:
            if conn.synapse is not None:
                if compat.is_neurons(conn.pre_obj) or (compat.is_neurons(conn.post_obj) and not _ens_to_neurons(conn)):
                    raise NotImplementedError("filtered neuron-to-neuron connections are outside the hot path")
                self.filt_col[conn] = alloc("filt", conn, self._out_size(conn))
        for probe in self.probes:
            if compat.kind(probe.obj) in ("node", "ensemble") and probe.synapse is not None:
                self.filt_col[probe] = alloc("filt", probe, probe.size_in)
        # decoded outputs, grouped per ensemble so that small ensembles own one contiguous slot
        for ens in self.ensembles:
            for c in self.ens_dec_conns[ens]:
                self.dec_col[c] = alloc("dec", c, self._out_size(c))
        for node in self.nodes:
            if self.node_kind[node] == "fn":
                self.fn_col[node] = alloc("fn", node, node.size_out)

    def _eye(self, c0, size):
        return sp.csr_matrix((np.ones(size), (np.arange(size), c0 + np.arange(size))), shape=(size, self.ncol))

    # ------------------------------------------------------------------ expressions
    def expr_out(self, node):
        memo = self._memo_out
        if node in memo:
            if memo[node] is None:
                raise RuntimeError(f"algebraic loop through pass-through node {node!r} (no synapse in the cycle)")
            return memo[node]
        memo[node] = None
        kind = self.node_kind[node]
        if kind == "table":
            m = self._eye(self.tab_col[node], node.size_out)
        elif kind == "const":
            vals = np.asarray(node.output, dtype=np.float64)
            m = sp.csr_matrix((vals, (np.arange(vals.size), np.zeros(vals.size, dtype=int))),
                              shape=(vals.size, self.ncol))
        elif kind == "fn":
            m = self._eye(self.fn_col[node], node.size_out)
        else:
            m = self.expr_in(node)
        memo[node] = m
        return m

    def expr_in(self, obj):
        size = obj.size_in
        total = None
        for conn in self.incoming.get(obj, []):
            term = _place(_idx(conn.post_slice, size), size, self.conn_value(conn))
            total = term if total is None else total + term
        return sp.csr_matrix((size, self.ncol)) if total is None else total.tocsr()

    def conn_value(self, conn):
        if conn.synapse is not None:
            return self._eye(self.filt_col[conn], conn.size_out)
        return self.weighted(conn)

    def weighted(self, conn):
        pre = conn.pre_obj
        if compat.is_ensemble(pre):
            return self._dec_expr(conn)
        if compat.is_neurons(pre):
            raise NotImplementedError("connections from ens.neurons are outside the hot path")
        src = self.expr_out(pre)
        rows = _idx(conn.pre_slice, pre.size_out)
        if not (len(rows) == src.shape[0] and np.array_equal(rows, np.arange(src.shape[0]))):
            src = src[rows]
        return _apply_transform(compat.transform_of(conn), src).tocsr()

    # ------------------------------------------------------------------ main
    def lower(self, chunk_cap):
        plan = self.plan
        self.classify_nodes()
        self.classify_ensembles()
        self.enumerate_sources()
        self._memo_out = {}
        dt = self.dt

        # ---- sinks
        ens_in, ens_jn, voja_rule, pes_rule = {}, {}, {}, {}
        for ens in self.ensembles:
            ens_in[ens] = self.expr_in(ens)
            jn = []
            for conn in self.incoming.get(ens.neurons, []):
                tr = compat.transform_of(conn)
                if tr is None or np.ndim(tr) != 2 or conn.post_slice != slice(None):
                    raise NotImplementedError("neuron-direct connections need a full (n x m) transform")
                pre = conn.pre_obj
                if _ens_to_neurons(conn):
                    if conn.synapse is None:
                        raise NotImplementedError("unfiltered ensemble -> neurons connections are outside the hot path")
                    u = self._eye(self.filt_col[conn], conn.size_mid)
                else:
                    u = self.expr_out(pre)[_idx(conn.pre_slice, pre.size_out)]
                G = self.model.params[ens].gain[:, None] * tr
                jn.append((u.tocsr(), G))
            if jn:
                ens_jn[ens] = (sp.vstack([u for u, _ in jn]).tocsr(), np.hstack([G for _, G in jn]))
        for conn in self.conns:
            rule = conn.learning_rule
            if rule is None:
                continue
            lrt = rule.learning_rule_type
            rin = self.expr_in(rule)
            if compat.rule_kind(lrt) == "voja":
                if lrt.post_synapse is not None:
                    raise NotImplementedError("Voja with a post_synapse is outside the hot path")
                one = sp.csr_matrix(([1.0], ([0], [0])), shape=(1, self.ncol))
                voja_rule[conn.post_obj] = (conn, (one + rin).tocsr(), lrt)
            elif compat.rule_kind(lrt) == "pes":
                pes_rule[conn] = (rin, lrt)
            else:
                raise NotImplementedError(type(lrt).__name__)
        fn_in = {n: self.expr_in(n) for n in self.nodes if self.node_kind[n] == "fn"}
        filt_in = {}
        for key in self.filt_col:
            if compat.is_connection(key):
                filt_in[key] = self.weighted(key)
            else:  # probe with synapse
                filt_in[key] = self._probe_expr(key)

        # ---- levels
        INF = 1 << 20
        col_level = np.zeros(self.ncol, dtype=np.int64)
        pending_ens, pending_fn = set(self.ensembles), set(fn_in)
        dec_width = {c: self._out_size(c) for c in self.dec_col}
        for c in pes_rule:
            col_level[self.dec_col[c]:self.dec_col[c] + dec_width[c]] = INF
        unresolved = np.zeros(self.ncol, dtype=bool)
        for c, col0 in self.dec_col.items():
            size = dec_width[c]
            if not (compat.is_connection(c) and c in pes_rule):
                unresolved[col0:col0 + size] = True
        for n, col0 in self.fn_col.items():
            unresolved[col0:col0 + n.size_out] = True

        def cols_of(*mats):
            return np.unique(np.concatenate([m.indices for m in mats if m is not None] + [np.zeros(0, dtype=np.int64)]))

        ens_level, fn_level = {}, {}
        progress = True
        while (pending_ens or pending_fn) and progress:
            progress = False
            for ens in list(pending_ens):
                mats = [ens_in[ens]]
                if ens in ens_jn:
                    mats.append(ens_jn[ens][0])
                if ens in voja_rule:
                    mats.append(voja_rule[ens][1])
                cols = cols_of(*mats).astype(np.int64)
                if cols.size and unresolved[cols].any():
                    continue
                lvl = int(col_level[cols].max()) if cols.size else 0
                if lvl >= INF:
                    raise NotImplementedError("a PES-learned connection feeds an ensemble without a synapse")
                ens_level[ens] = lvl
                for c in self.ens_dec_conns[ens]:
                    if compat.is_connection(c) and c in pes_rule:
                        continue
                    size = dec_width[c]
                    col_level[self.dec_col[c]:self.dec_col[c] + size] = lvl + 1
                    unresolved[self.dec_col[c]:self.dec_col[c] + size] = False
                pending_ens.discard(ens)
                progress = True
            for node in list(pending_fn):
                cols = cols_of(fn_in[node]).astype(np.int64)
                if cols.size and unresolved[cols].any():
                    continue
                lvl = int(col_level[cols].max()) if cols.size else 0
                if lvl >= INF:
                    raise NotImplementedError("a PES-learned connection feeds a function node without a synapse")
                fn_level[node] = lvl
                col_level[self.fn_col[node]:self.fn_col[node] + node.size_out] = lvl + 1
                unresolved[self.fn_col[node]:self.fn_col[node] + node.size_out] = False
                pending_fn.discard(node)
                progress = True
        if pending_ens or pending_fn:
            raise RuntimeError("same-step dependency cycle between ensembles / function nodes")
        n_levels = 1 + max(list(ens_level.values()) + list(fn_level.values()) + [0])

        # ---- exact same-step dependencies of every level: which producer kinds of which earlier level its sink rows read
        #      (bit 0: decoded outputs of narrow ensembles, 1: static decoders of wide ensembles, 2: grid clean-up nodes,
        #       3: gate nodes).  The launch sequence lets a level start as soon as THOSE producers are done instead of
        #      waiting for every kernel of the previous level.
        col_prod = np.zeros(self.ncol, dtype=np.int64)          # producer kind bit of a column (0: exists at step start)
        col_plvl = np.zeros(self.ncol, dtype=np.int64)          # level of its producer
        for c, col0 in self.dec_col.items():
            if compat.is_connection(c) and c in pes_rule:
                continue
            ens_c = c.pre_obj if compat.is_connection(c) else c.obj
            col_prod[col0:col0 + dec_width[c]] = 1 if self.is_small[ens_c] else 2
            col_plvl[col0:col0 + dec_width[c]] = ens_level[ens_c]
        for node, col0 in self.fn_col.items():
            col_prod[col0:col0 + node.size_out] = 4 if self.node_op[node].kind == "cleanup" else 8
            col_plvl[col0:col0 + node.size_out] = fn_level[node]
        level_deps = np.zeros((n_levels, n_levels), dtype=np.int32)

        def note_deps(lvl, *mats):
            cols = cols_of(*mats).astype(np.int64)
            cols = cols[col_prod[cols] != 0]
            if cols.size:
                np.bitwise_or.at(level_deps[lvl], col_plvl[cols], col_prod[cols].astype(np.int32))
        for ens in self.ensembles:
            mats = [ens_in[ens]]
            if ens in ens_jn:
                mats.append(ens_jn[ens][0])
            if ens in voja_rule:
                mats.append(voja_rule[ens][1])
            note_deps(ens_level[ens], *mats)
        for node, mat in fn_in.items():
            note_deps(fn_level[node], mat)
        plan.arrays["level_deps"] = level_deps

        # ---- device column map: row 0 ones | filters A | filters B | this step's table rows | scratch
        NF = sum(1 for k in self.col_kind if k == "filt")
        NT = sum(1 for k in self.col_kind if k == "tab")
        dev_col = np.zeros(self.ncol, dtype=np.int64)
        next_filt, next_tab = 1, 0
        tab_row0 = 1 + 2 * NF
        next_scratch = tab_row0 + 2 * NT      # two copies of the input rows (step parity), so step t+1's inputs can be
                                             # written while step t still reads its own
        for c in range(1, self.ncol):
            k = self.col_kind[c]
            if k == "filt":
                dev_col[c] = next_filt
                next_filt += 1
            elif k == "tab":
                dev_col[c] = tab_row0 + next_tab
                next_tab += 1
            else:
                dev_col[c] = next_scratch
                next_scratch += 1
        NV = next_scratch
        # logical columns ncol + k (k-th table column): the same input row of the NEXT step, i.e. the other parity's copy -
        # used by the fused end-of-step rows that already evaluate the next step's level-0 sink rows (see below)
        dev_col = np.concatenate([dev_col, TABN_BASE + np.arange(NT, dtype=np.int64)])
        tab_ord = {c: k for k, c in enumerate(c for c in range(self.ncol) if self.col_kind[c] == "tab")}
        self.dev_col = dev_col
        for node, c0 in self.tab_col.items():
            plan.tables.append((node, int(dev_col[c0] - tab_row0), node.size_out))
        for key, c0 in self.filt_col.items():
            size = self._out_size(key)
            plan.filters[key] = (int(dev_col[c0] - 1), size)

        # ---- CSR program
        csr_ptr, csr_idx, csr_val = [0], [], []   # csr_idx holds parity-0 vec rows (filter columns in half A)

        def add_rows(mat):
            mat = mat.tocsr()
            mat.sum_duplicates()
            mat.eliminate_zeros()
            row0 = len(csr_ptr) - 1
            for r in range(mat.shape[0]):
                lo, hi = mat.indptr[r], mat.indptr[r + 1]
                cols = mat.indices[lo:hi]
                order = np.argsort(cols, kind="stable")
                csr_idx.extend(dev_col[cols[order]].tolist())
                csr_val.extend(mat.data[lo:hi][order].tolist())
                pad = (-(hi - lo)) % CSR_PAD      # the kernels walk rows 8 entries at a time (no tail)
                csr_idx.extend([0] * pad)
                csr_val.extend([0.0] * pad)
                csr_ptr.append(len(csr_idx))
            return row0

        # ---- sink rows are materialised once per level by k_lin-style passes (kinds 3 / 4) into vec scratch;
        #      the consumers' descriptors carry the vec row of their input
        mat_rows = [[] for _ in range(n_levels)]
        mat_log = []          # (matrix, first vec row, level, previous_view) of every materialised block

        def materialize(mat, lvl, previous_view=False):
            nonlocal NV
            r0, nrow = add_rows(mat), mat.shape[0]
            v0 = NV
            NV += nrow
            for i in range(nrow):
                mat_rows[lvl].append([r0 + i, 4 if previous_view else 3, v0 + i])
            mat_log.append((mat.tocsr(), v0, lvl, previous_view))
            return v0

        # ---- weights + descriptors
        W = []            # static float32 weights (shared by all trials)
        w_len = 0

        def add_w(arr, align=4):
            nonlocal w_len
            pad = (-w_len) % align
            if pad:
                W.append(np.zeros(pad, dtype=np.float32))
                w_len += pad
            off = w_len
            a = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
            W.append(a)
            w_len += a.size
            return off

        # per-trial plans: the seed-dependent scalars / narrow-ensemble rows live in the per-trial weight arena instead;
        # their offsets come from _pt_blocks (the same walk fills the arena for every trial's model)
        pt_off = {}
        if self.per_trial:
            off = 0
            for key, a in self._pt_blocks(self.model):
                off += (-off) % 4
                pt_off[key] = off
                off += a.size

        ntypes, ntype_ids = [], {}

        def ntype_id(nt):
            if compat.neuron_kind(nt) == "lif":
                if nt.min_voltage != 0:
                    raise NotImplementedError("the packed one-word LIF state needs min_voltage == 0 (nengo's default)")
                # polynomial expm1 / log1p are exact to fp32 for dt / tau_rc <= 1/16 (SSB kernels header)
                fast = 1.0 if dt / nt.tau_rc <= 0.0625 else 0.0
                key = (NT_LIF, nt.tau_rc, nt.tau_ref, nt.min_voltage, nt.amplitude, fast, 0.0, 0.0)
            elif compat.neuron_kind(nt) == "lifrate":
                key = (NT_LIFRATE, nt.tau_rc, nt.tau_ref, 0.0, nt.amplitude, 0.0, 0.0, 0.0)
            elif compat.neuron_kind(nt) == "relu":
                key = (NT_RELU, 0.0, 0.0, 0.0, nt.amplitude, 0.0, 0.0, 0.0)
            else:
                raise NotImplementedError(f"neuron type {type(nt).__name__} is not supported by the B200 backend")
            if key not in ntype_ids:
                ntype_ids[key] = len(ntypes)
                ntypes.append(key)
            return ntype_ids[key]

        small_desc = [[] for _ in range(n_levels)]
        big_desc = [[] for _ in range(n_levels)]
        dec_desc = [[] for _ in range(n_levels)]
        pes_desc, cleanup_desc, gate_desc = [], [[] for _ in range(n_levels)], [[] for _ in range(n_levels)]
        pes_trace, cleanup_s64 = [], [[] for _ in range(n_levels)]
        nn = n_act = n_lenc = n_ldec = n_afilt = n_ldec_words = 0
        n_small = n_big = 0
        n_pt_enc = n_pt_dec = 0
        n_part = n_jtiles = 0
        pes_level = -1

        def splitk(c, size_out):
            """(n_chunks, part_off, counter0) of a decode / PES item."""
            nonlocal n_part, n_jtiles
            k = self.dec_chunks[c]
            part_off, counter0 = n_part, n_jtiles
            if k > 1:
                n_part += k * size_out
            n_jtiles += -(-size_out // DEC_TILE)
            return [k, part_off, counter0]

        for ens in self.ensembles:
            p = self.model.params[ens]
            n, dims = ens.n_neurons, ens.dimensions
            lvl = ens_level[ens]
            outs = self.ens_dec_conns[ens]
            nout = sum(self._out_size(c) for c in outs)
            is_small = self.is_small[ens]
            state0 = nn
            nn += n
            plan.ens_state[ens] = (state0, n)
            tid = ntype_id(ens.neuron_type)
            in_row0 = materialize(ens_in[ens], lvl)
            if is_small:
                n_small += 1
                decs = [self._dec_weights(c) for c in outs]  # each (size_out x n)
                stride = 1 + dims + nout
                stride += (-stride) % 4
                packed = np.zeros((n, stride))
                packed[:, 0] = p.bias
                packed[:, 1:1 + dims] = p.scaled_encoders
                if decs:
                    packed[:, 1 + dims:1 + dims + nout] = np.vstack(decs).T
                w_off = pt_off[(ens, "packed")] if self.per_trial else add_w(packed)
                out_vec = int(dev_col[self.dec_col[outs[0]]]) if outs else 0
                # decoded slots of one ensemble are contiguous by construction
                small_desc[lvl].append([n, dims, nout, state0, w_off, in_row0, out_vec, tid, stride])
                continue

            n_big += 1
            act0 = n_act
            n_act += n
            plan.ens_act[ens] = act0
            dpad = dims + ((-dims) % 4)
            flags = 0
            voja_alpha = 0.0
            voja_row = scale_off = 0
            if ens in voja_rule:
                conn, lrow, lrt = voja_rule[ens]
                flags |= 1
                enc_off = n_lenc
                n_lenc += n * dims
                plan.learned_enc[ens] = (enc_off, n, dims)
                voja_alpha = lrt.learning_rate * dt
                voja_row = materialize(lrow, lvl)
                scale_off = pt_off[(ens, "scale")] if self.per_trial else add_w(p.gain / ens.radius)
            elif self.per_trial:
                # static encoders of a trial's own model: per-trial rows of the lenc arena, walked by the Voja kernel with
                # alpha = 0 (flags bit 0), bias from the per-trial weight arena (flags bit 2)
                flags |= 1
                enc_off = n_lenc
                n_lenc += n * dims
                n_pt_enc += n * dims
                plan.pt_enc[ens] = (enc_off, n, dims)
            else:
                enc = np.zeros((n, dpad))
                enc[:, :dims] = p.scaled_encoders
                enc_off = add_w(enc)
            if self.per_trial:
                flags |= 4
                bias_off = pt_off[(ens, "bias")]
            else:
                bias_off = add_w(p.bias)
            jn_row0 = jn_m = jn_w = 0
            if ens in ens_jn:
                u, G = ens_jn[ens]
                jn_row0, jn_m = materialize(u, lvl), u.shape[0]
                jn_w = pt_off[(ens, "jn")] if self.per_trial else add_w(G)
                flags |= 2
            big_desc[lvl].append([n, dims, dpad, state0, act0, enc_off, bias_off, in_row0, tid, flags,
                                  jn_row0, jn_m, jn_w, voja_row, scale_off,
                                  int(np.float32(voja_alpha).view(np.int32))])
            for c in outs:
                size_out = self._out_size(c)
                out_vec = int(dev_col[self.dec_col[c]])
                if compat.is_connection(c) and c in pes_rule:
                    rin, lrt = pes_rule[c]
                    d_off = n_ldec
                    # device layout of a learned decoder (csrc/ssb_pes.cuh): per trial group [neuron][trial][JP] floats,
                    # JP = size_out rounded up to 4 -> JP arena rows of 32 floats per neuron
                    n_ldec += (-(-size_out // 4) * 4) * n
                    n_ldec_words += size_out * n
                    plan.learned_dec[c] = (d_off, size_out, n)
                    a_off = n_afilt
                    n_afilt += n
                    kinds = {self.col_kind[cc] for cc in np.unique(rin.indices)}
                    if not kinds <= {"filt", "const"}:
                        raise NotImplementedError(
                            "PES error input must arrive through a synapse (the delta of step t-1 is rebuilt "
                            "from the previous filter values)")
                    err_row0 = materialize(rin, 0, previous_view=True)   # filter / constant columns only
                    alpha = -lrt.learning_rate * dt / n
                    if lrt.pre_synapse is None:
                        decay = 0.0
                    else:
                        decay = float(np.exp(-dt / compat.synapse_tau(lrt.pre_synapse)))
                    pes_trace.append((act0, a_off, n, np.float32(decay), np.float32(1.0 - decay)))
                    pes_level = max(pes_level, lvl)
                    pes_desc.append([n, size_out, d_off, a_off, act0, err_row0, out_vec,
                                     int(np.float32(alpha).view(np.int32)),
                                     int(np.float32(decay).view(np.int32)),
                                     int(np.float32(1.0 - decay).view(np.int32))] + splitk(c, size_out))
                elif self.per_trial:
                    jpad = size_out + ((-size_out) % DEC_TILE)
                    d_off = n_ldec
                    n_ldec += (-(-size_out // 4) * 4) * n            # [neuron][trial][JP], as the learned decoders
                    n_pt_dec += size_out * n
                    plan.pt_dec[c] = (d_off, size_out, n)
                    dec_desc[lvl].append([n, size_out, jpad, act0, d_off, out_vec] + splitk(c, size_out))
                else:
                    jpad = size_out + ((-size_out) % DEC_TILE)
                    Wd = np.zeros((n, jpad))
                    Wd[:, :size_out] = self._dec_weights(c).T
                    w_off = add_w(Wd)
                    plan.static_dec[c] = (w_off, size_out, jpad, n)
                    dec_desc[lvl].append([n, size_out, jpad, act0, w_off, out_vec] + splitk(c, size_out))

        for node, mat in fn_in.items():
            op = self.node_op[node]
            lvl = fn_level[node]
            in_row0 = materialize(mat, lvl)
            out_vec = int(dev_col[self.fn_col[node]])
            if op.kind == "cleanup":
                S = op.sample_ssps
                G, d = S.shape
                dpad = d + ((-d) % 4)
                Sp = np.zeros((G, dpad))
                Sp[:, :d] = S
                s_off = add_w(Sp)
                cleanup_desc[lvl].append([G, d, dpad, s_off, in_row0, out_vec])
                cleanup_s64[lvl].append(np.ascontiguousarray(S, dtype=np.float64).reshape(-1))
            else:
                gate_desc[lvl].append([op.d, in_row0, out_vec,
                                       int(np.float32(op.shift_rate).view(np.int32)),
                                       int(np.float32(op.update_thres).view(np.int32)),
                                       int(np.float32(op.atol).view(np.int32))])

        # ---- final stage rows: filters then probes.  A filter row that reads only step-start columns and decoded outputs of
        #      LEVEL-0 narrow ensembles (in SLAM: every VCO filter) is "early": it can run as soon as that kernel is done,
        #      next to the other chains of the step, instead of in the end-of-step launch that waits for everything.  (It
        #      writes the filter half nobody reads in this step; the only readers of that half - the previous-view rows of
        #      the PES error - are materialised by the very first launch of the step.)
        def early_rows(mat):
            ok = np.ones(mat.shape[0], dtype=bool)
            late_col = ~((col_prod == 0) | ((col_prod == 1) & (col_plvl == 0)))
            late_col |= col_level >= INF                     # PES-decoded columns
            m = mat.tocsr()
            for r in range(m.shape[0]):
                cols = m.indices[m.indptr[r]:m.indptr[r + 1]]
                ok[r] = not late_col[cols].any()
            return ok

        lin_rows, lin_ab = [], []
        lin_early, lin_early_ab = [], []
        filt_update = []      # (f0, size, a, b, U) of every Lowpass: y_new = a * y_old + b * (U . columns)
        for key, mat in filt_in.items():
            f0, size = plan.filters[key]
            tau = compat.synapse_tau(key.synapse)
            a64 = np.exp(-dt / tau) if tau > 0 else 0.0     # Lowpass(0) (slam_loihi.py:233): a pure one-step delay
            a, b = np.float32(a64), np.float32(1.0 - a64)
            filt_update.append((f0, size, float(a), float(b), mat.tocsr()))
            r0 = add_rows(mat)
            if mat.shape[0] != size:
                raise AssertionError("filter size mismatch")
            # (single-level plans - PathIntegration - have no other chain to overlap with: the split would only add a launch;
            #  measured on B200: SLAM 264.9 -> 260.3 us per step with the split, PathIntegration d = 97 74.0 -> 78.0 us)
            early = (early_rows(mat) if n_levels > 1 and os.environ.get("SSB_LIN_EARLY", "1") != "0"
                     else np.zeros(size, bool))
            for i in range(size):
                if early[i]:
                    lin_early.append([r0 + i, 0, f0 + i])
                    lin_early_ab.append([a, b])
                else:
                    lin_rows.append([r0 + i, 0, f0 + i])
                    lin_ab.append([a, b])
        for act0, a_off, n, a, b in pes_trace:  # PES pre-synaptic activity traces (Lowpass of the spikes)
            for i in range(n):
                lin_rows.append([act0 + i, 2, a_off + i])
                lin_ab.append([a, b])
        n_probe_rows = 0
        for probe in self.probes:
            period = 1 if probe.sample_every is None else int(round(probe.sample_every / dt))
            obj = probe.obj
            if compat.is_neurons(obj):
                # neuron-output probes (run_pathint_gif.py:157-159: ``ea_ensembles[k].neurons[:500]``, synapse=None):
                # the step's activity rows are copied to the probe block by the row program (kind 5)
                if probe.synapse is not None:
                    raise NotImplementedError("filtered probes on ens.neurons are outside the hot path")
                if obj.ensemble not in plan.ens_act:
                    raise NotImplementedError(f"probe {probe!r}: the ensemble is not simulated on the device")
                sel = _idx(probe.slice, obj.size_out)
                info = ProbeInfo(probe, "rows", n_probe_rows, len(sel), period)
                for i, k in enumerate(sel):
                    lin_rows.append([plan.ens_act[obj.ensemble] + int(k), 5, n_probe_rows + i])
                    lin_ab.append([0.0, 1.0])
                n_probe_rows += len(sel)
            elif compat.kind(obj) in ("node", "ensemble"):
                if probe in self.filt_col:
                    mat = self._eye(self.filt_col[probe], probe.size_in)
                else:
                    mat = self._probe_expr(probe)
                r0 = add_rows(mat)
                info = ProbeInfo(probe, "rows", n_probe_rows, probe.size_in, period)
                for i in range(probe.size_in):
                    lin_rows.append([r0 + i, 1, n_probe_rows + i])
                    lin_ab.append([0.0, 1.0])
                n_probe_rows += probe.size_in
            elif compat.is_connection(obj) and probe.attr == "weights":
                if obj not in plan.learned_dec:
                    raise NotImplementedError("'weights' probes are supported on PES-learned connections")
                info = ProbeInfo(probe, "weights", period=period, conn=obj)
            elif compat.is_learning_rule(obj) and probe.attr == "scaled_encoders":
                info = ProbeInfo(probe, "scaled_encoders", period=period, ens=obj.connection.post_obj)
            else:
                raise NotImplementedError(f"probe {probe!r} is outside the hot path")
            plan.probes.append(info)

        # every probe row is written each step; make sure late (PES) columns are only used there
        # ---- assemble
        def arr(rows, width):
            return np.asarray(rows, dtype=np.int32).reshape(-1, width)

        # ---- step fusion: the level-0 sink rows of step t+1 read only constants, filter states and input tables, and the
        #      new filter states are themselves linear in what step t has at its end (y_new = a y_old + b U v).  Composing
        #      the two gives rows over step t's columns (+ the next step's table rows) that the END-OF-STEP launch of step t
        #      can evaluate next to the filter updates - so the first launch of step t+1, on which every chain of the step
        #      waits, disappears from the critical path.  Rows on the previous step's view (kind 4: the PES error) cannot be
        #      moved; they stay in a small residual launch that only the PES chain waits for.
        lin_fused = []
        n_res = 0
        # Measured on B200 and NOT the default (SSB_LIN_FUSE=1 enables it): the composed rows are 2.7x denser than level 0's
        # own (configs[1]: 66 808 vs 24 648 entries), the fused launch takes as long as the two launches it replaces and the
        # step gets slower - 273.0 vs 257.8 us (configs[1]), 75.1 vs 73.5 (PathIntegration d = 97), 389 vs 377 (SLAMView d = 97);
        # profiles/r02h_perf_step_fusion.log.  A k_lin launch costs its latency chain, not its arithmetic.
        if n_levels >= 1 and NF > 0 and os.environ.get("SSB_LIN_FUSE", "0") == "1":
            ncol, ncx = self.ncol, self.ncol + NT
            filt_cols = np.array([c for c in range(ncol) if self.col_kind[c] == "filt"], dtype=np.int64)
            f_index = dev_col[filt_cols] - 1                                   # filter index of a logical filter column
            rows_n, cols_n, vals_n = [], [], []
            f_has = np.zeros(NF, dtype=bool)
            for f0, size, a, b, U in filt_update:
                U = U.tocoo()
                rows_n += (f0 + U.row).tolist()
                cols_n += U.col.tolist()
                vals_n += (b * U.data).tolist()
                f_has[f0:f0 + size] = True
            # a * y_old: the logical column of filter f
            col_of_f = np.zeros(NF, dtype=np.int64)
            col_of_f[f_index] = filt_cols
            a_of_f = np.zeros(NF)
            for f0, size, a, b, U in filt_update:
                a_of_f[f0:f0 + size] = a
            rows_n += list(range(NF))
            cols_n += col_of_f.tolist()
            vals_n += a_of_f.tolist()
            N = sp.csr_matrix((vals_n, (rows_n, cols_n)), shape=(NF, ncx))      # new filter states from step-t columns
            sel_f = sp.csr_matrix((np.ones(len(filt_cols)), (filt_cols, f_index)), shape=(ncol, NF))
            keep = np.array([self.col_kind[c] == "const" for c in range(ncol)])
            tabs = np.array([c for c in range(ncol) if self.col_kind[c] == "tab"], dtype=np.int64)
            sel_c = sp.csr_matrix((np.ones(int(keep.sum())), (np.flatnonzero(keep), np.flatnonzero(keep))), shape=(ncol, ncx))
            sel_t = sp.csr_matrix((np.ones(len(tabs)), (tabs, ncol + np.array([tab_ord[c] for c in tabs], dtype=np.int64))),
                                  shape=(ncol, ncx))
            ok = bool(f_has.all())
            blocks, nnz_in, nnz_out = [], 0, 0
            for M, v0, lvl, prev in mat_log:
                if lvl != 0 or prev:
                    continue
                kinds = {self.col_kind[c] for c in np.unique(M.indices)}
                if not kinds <= {"const", "filt", "tab"}:
                    ok = False
                    break
                comp = (M @ sel_f @ N + M @ sel_c + M @ sel_t).tocsr()
                blocks.append((comp, v0))
                nnz_in += M.nnz
                nnz_out += comp.nnz
            # fusion removes a LATENCY-bound launch; it composes matrices, so it must stay small: at d = 649 the level-0
            # rows are 1 300 x 649 GEMMs (throughput-bound launches) and composing them with the inverse DFT doubles the work
            plan.stats_fusion = dict(nnz_level0=int(nnz_in), nnz_filters=int(N.nnz), nnz_fused=int(nnz_out))
            if ok and blocks and nnz_out <= 4 * (nnz_in + N.nnz) and nnz_out <= FUSE_MAX_NNZ:
                for comp, v0 in blocks:
                    r0 = add_rows(comp)
                    for i in range(comp.shape[0]):
                        lin_fused.append([r0 + i, 3, v0 + i])
                # the residual launch of a fused step: level 0's previous-view rows, kept at the end of its segment
                mat_rows[0].sort(key=lambda r: r[1] == 4)
                n_res = sum(1 for r in mat_rows[0] if r[1] == 4)

        lin_final = lin_rows
        lin_final_ab = lin_ab
        lin_rows, lin_ab, level_rows = [], [], []
        for lvl in range(n_levels):
            level_rows.append((len(lin_rows), len(mat_rows[lvl])))
            lin_rows += mat_rows[lvl]
            lin_ab += [[0.0, 1.0]] * len(mat_rows[lvl])
        lin0 = len(lin_rows)
        lin_rows += lin_early + lin_final + lin_fused      # final rows: [early | late | next step's level-0 rows (fused)]
        lin_ab += lin_early_ab + lin_final_ab + [[0.0, 1.0]] * len(lin_fused)
        stages = []  # per level counts/offsets into the concatenated descriptor arrays
        cat = {k: [] for k in ("small", "big", "dec", "cleanup", "gate")}
        for lvl in range(n_levels):
            # heavy-first ordering inside a launch evens out the tail
            small_desc[lvl].sort(key=lambda r: (-r[0], -(r[1] + r[2])))
            entry = []
            for name, lst in (("small", small_desc[lvl]), ("big", big_desc[lvl]), ("dec", dec_desc[lvl]),
                              ("cleanup", cleanup_desc[lvl]), ("gate", gate_desc[lvl])):
                entry += [len(cat[name]), len(lst)]
                cat[name] += lst
            stages.append(entry + list(level_rows[lvl]))
        plan.arrays.update({
            "csr_ptr": np.asarray(csr_ptr, dtype=np.int32),
            "csr_ent0": self._entries(csr_idx, csr_val, 0, NF, tab_row0, NT),
            "csr_ent1": self._entries(csr_idx, csr_val, NF, NF, tab_row0, NT),
            # 8 floats of slack: bulk copies of bias / current weights round their length up to 16 bytes
            "weights": np.concatenate(W + [np.zeros(8, dtype=np.float32)]),
            **({"weights_pt": self.trial_weights(self.model)} if self.per_trial else {}),
            "ens_small": arr(cat["small"], 9),
            "ens_big": arr(cat["big"], 16),
            "dec": arr(cat["dec"], 9),
            "pes": arr(pes_desc, 13),
            "cleanup": arr(cat["cleanup"], 6),
            "gate": arr(cat["gate"], 6),
            "lin_rows": arr(lin_rows, 3),
            "lin_ab": np.asarray(lin_ab, dtype=np.float32).reshape(-1, 2),
            "stages": arr(stages, 12),
            "ntypes": np.asarray(ntypes, dtype=np.float32).reshape(-1, 8),
            "cleanup_s64": np.concatenate([a for lvl in cleanup_s64 for a in lvl] + [np.zeros(0)]),
        })
        plan.cleanup_nodes = [n for lvl in range(n_levels) for n in fn_in
                              if self.node_op[n].kind == "cleanup" and fn_level[n] == lvl]
        plan.scalars.update(dict(dt=dt, nv=NV, nf=NF, nt=NT, tab_row0=tab_row0, nn=nn, n_act=n_act, n_lenc=n_lenc,
                                 n_ldec=n_ldec, n_afilt=n_afilt, n_probe=n_probe_rows, n_levels=n_levels,
                                 chunk_cap=chunk_cap, n_part=n_part, n_jtiles=n_jtiles, pes_level=pes_level,
                                 lin0=lin0, n_lin=len(lin_rows) - lin0 - len(lin_fused), n_lin_early=len(lin_early),
                                 n_lin_fused=len(lin_fused), n_lvl0_res=n_res))
        n_static = int(sum(a.size for a in W))
        if self.per_trial:
            plan.scalars["per_trial_weights"] = 1.0
            # static weights a trial owns (read once per step, never written): narrow rows + bias / scale + wide enc / dec
            n_pt_weights = int(plan.arrays["weights_pt"].size - 8) + n_pt_enc + n_pt_dec
        else:
            n_pt_weights = 0
        plan.stats = dict(n_pt_weights=n_pt_weights, n_neurons=nn, n_filter_states=NF, n_learned=n_lenc - n_pt_enc + n_ldec_words,
                          n_static_weights=n_static,
                          n_table_words=NT, n_probe_words=n_probe_rows, n_small=n_small, n_big=n_big,
                          n_levels=n_levels, csr_nnz=len(csr_idx), n_afilt=n_afilt, n_act=n_act)
        n_small_neurons = int(sum(e.n_neurons for e in self.ensembles if self.is_small[e]))
        n_voja_neurons = int(sum(e.n_neurons for e in voja_rule))
        # SURVEY.md §8(d) traffic model split by the kernel that owns each stream (bytes per trial-step)
        plan.stats["n_small_neurons"] = n_small_neurons
        plan.stats["bytes_by_kind"] = {
            "ens_small": 16 * n_small_neurons,
            "ens_wide": 16 * (nn - n_small_neurons - n_voja_neurons),
            "ens_voja": 16 * n_voja_neurons + 8 * (n_lenc - n_pt_enc),
            "pes": 8 * n_ldec_words,
            "lin": 8 * (NF + n_afilt) + 4 * n_probe_rows,
            "inputs": 4 * NT,
        }
        return plan

    def _pt_blocks(self, model):
        """(key, float array) blocks of the per-trial weight arena in arena order: a narrow ensemble's packed
        ``[bias | scaled encoders | decoders]`` rows; a wide ensemble's Voja scale (if learned), bias and direct neuron-current
        weights."""
        voja_posts = {c.post_obj for c in self.conns
                      if c.learning_rule is not None and compat.rule_kind(c.learning_rule.learning_rule_type) == "voja"}
        saved, self.model = self.model, model
        try:
            for ens in self.ensembles:
                p = model.params[ens]
                if self.is_small[ens]:
                    outs = self.ens_dec_conns[ens]
                    nout = sum(self._out_size(c) for c in outs)
                    dims = ens.dimensions
                    stride = 1 + dims + nout
                    stride += (-stride) % 4
                    packed = np.zeros((ens.n_neurons, stride))
                    packed[:, 0] = p.bias
                    packed[:, 1:1 + dims] = p.scaled_encoders
                    if outs:
                        packed[:, 1 + dims:1 + dims + nout] = np.vstack([self._dec_weights(c) for c in outs]).T
                    yield (ens, "packed"), np.ascontiguousarray(packed, dtype=np.float32).reshape(-1)
                else:
                    if ens in voja_posts:
                        yield (ens, "scale"), np.asarray(p.gain / ens.radius, dtype=np.float32).reshape(-1)
                    yield (ens, "bias"), np.asarray(p.bias, dtype=np.float32).reshape(-1)
                    trs = [compat.transform_of(c) for c in self.incoming.get(ens.neurons, [])]
                    if trs:      # direct neuron currents: gain * transform, [n][m] (see ens_jn in lower())
                        yield (ens, "jn"), np.asarray(p.gain[:, None] * np.hstack(trs), dtype=np.float32).reshape(-1)
        finally:
            self.model = saved

    def trial_weights(self, model):
        """The per-trial weight arena of ``model`` (one float per arena row), blocks aligned to 4 like the shared array."""
        out, off = [], 0
        for _, a in self._pt_blocks(model):
            pad = (-off) % 4
            if pad:
                out.append(np.zeros(pad, dtype=np.float32))
            out.append(a)
            off += pad + a.size
        return np.concatenate(out + [np.zeros(8, dtype=np.float32)])

    @staticmethod
    def _entries(idx, val, par, nf, tab_row0=0, nt=0):
        """CSR entries as (vec row, float32 coefficient bits) pairs with the filter columns resolved to
        the half that is read on steps of this parity (``par`` = 0 for even steps, ``nf`` for odd ones);
        input-table rows likewise point at the copy written for steps of this parity."""
        rows = np.asarray(idx, dtype=np.int64)
        nxt = rows >= TABN_BASE                              # table rows of the NEXT step: the other parity's copy
        if par and nt:
            rows = np.where((rows >= tab_row0) & (rows < tab_row0 + nt), rows + nt, rows)
        rows = np.where(nxt, rows - TABN_BASE + tab_row0 + (0 if par else nt), rows)
        rows = np.where((rows >= 1) & (rows <= nf), rows + par, rows)
        ent = np.empty((len(rows), 2), dtype=np.int32)
        ent[:, 0] = rows
        ent[:, 1] = np.asarray(val, dtype=np.float32).view(np.int32)
        return ent

    def _dec_weights(self, c):
        if compat.is_connection(c):
            if _ens_to_neurons(c):
                return np.asarray(self.model.params[c].decoders, dtype=np.float64)
            return np.asarray(self.model.params[c].weights, dtype=np.float64)
        return np.asarray(self.model.probe_conns[c], dtype=np.float64)

    def _probe_expr(self, probe):
        obj = probe.obj
        if compat.is_ensemble(obj):
            return self._dec_expr(probe)
        return self.expr_out(obj)[_idx(probe.slice, obj.size_out)].tocsr()


def lower(network, model: BuiltModel, chunk_cap=256, n_trials=1, per_trial=False) -> DevicePlan:
    """Network + built parameters -> :class:`DevicePlan` (``n_trials`` only tunes launch geometry).  ``per_trial``: every
    trial has its own network seed - seed-dependent static weights are laid out in per-trial arenas (``weights_pt``, wide
    encoders in ``lenc``, wide decoders in ``ldec``); ``model`` then only provides the layout and trial 0's values."""
    return _Lowerer(network, model, n_trials, per_trial).lower(chunk_cap)


def trial_weights(network, model: BuiltModel, n_trials=1):
    """Seed-dependent static weights of another built model of the same graph, for a plan lowered with ``per_trial=True``:
    ``(weights_pt array, {wide ens: scaled encoders [n, dims]}, {wide static decoder: weights [size_out, n]})``."""
    low = _Lowerer(network, model, n_trials, True)
    low.classify_ensembles()
    pes_conns = {c for c in low.conns
                 if c.learning_rule is not None and compat.rule_kind(c.learning_rule.learning_rule_type) == "pes"}
    enc, dec = {}, {}
    for ens in low.ensembles:
        if low.is_small[ens]:
            continue
        enc[ens] = np.asarray(model.params[ens].scaled_encoders, dtype=np.float32)
        for c in low.ens_dec_conns[ens]:
            if c not in pes_conns:
                dec[c] = np.asarray(low._dec_weights(c), dtype=np.float32)
    return low.trial_weights(model), enc, dec


def algorithmic_bytes_per_trial_step(stats, per_trial_weights=None):
    """SURVEY.md §8(d) traffic model (fp32): state R+W, filters R+W, learned R+W, inputs, probes; plans lowered with
    per-trial static weights also read every weight a trial owns once per step (``stats['n_pt_weights']``)."""
    b = 16 * stats["n_neurons"] + 8 * (stats["n_filter_states"] + stats["n_afilt"]) + 8 * stats["n_learned"]
    b += 4 * stats.get("n_pt_weights", 0)
    b += 4 * stats["n_table_words"] + 4 * stats["n_probe_words"]
    return int(b)
