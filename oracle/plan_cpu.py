"""TEST / BASELINE INFRASTRUCTURE (never imported by the package): an OPERATOR-MERGED CPU executor of a lowered
``DevicePlan`` for one trial.

``oracle/nengo_ref_sim.py`` restates nengo's reference simulator operator by operator (~5 500 NumPy calls per timestep
for SLAM d = 55) and is therefore several times slower than real nengo, whose optimiser merges operators over concatenated
signals.  This executor is the other end: the SAME arithmetic (float64 LIF / LIFRate / ReLU updates as
``nengo/neurons.py`` ``step``, Lowpass as zero-order hold, SimPES / SimVoja deltas applied at the next step) on the
algebraically collapsed plan, with every group of like operators done as ONE vector operation - all narrow ensembles of a
level in one concatenated update, each sink-row segment as one sparse mat-vec.  It is an upper bound of what operator
merging can reach on one core, so ``bench.py`` reports it next to the unmerged port as the honest lower companion of the
GPU / CPU ratio.  Semantics follow ``tests/plan_interp.py`` (the row-by-row interpreter of the same plan), against which
and against ``RefSimulator`` it is checked in ``tests/test_plan_cpu.py``.
"""
import numpy as np
import scipy.sparse as sp

NT_LIF, NT_LIFRATE, NT_RELU = 0, 1, 2


class MergedPlanSimulator:
    def __init__(self, plan, model, network, tables, trial_seed=None, dtype=np.float64):
        self.p, self.dt_ = plan, dtype
        a, sc = plan.arrays, plan.scalars
        if sc.get("per_trial_weights", 0):
            raise NotImplementedError("the merged CPU executor takes shared-weight plans")
        self.dt = float(sc["dt"])
        self.nf, self.nt, self.tab_row0 = int(sc["nf"]), int(sc["nt"]), int(sc["tab_row0"])
        self.nv = int(sc["nv"])
        self.vec = np.zeros(self.nv, dtype)
        self.vec[0] = 1
        nn = int(sc["nn"])
        self.v = np.zeros(nn, dtype)
        self.ref = np.zeros(nn, dtype)
        self.act = np.zeros(max(1, int(sc["n_act"])), dtype)
        self.lenc = np.zeros(max(1, int(sc["n_lenc"])), dtype)
        self.ldec = np.zeros(max(1, int(sc["n_ldec"])), dtype)
        self.afilt = np.zeros((2, max(1, int(sc["n_afilt"]))), dtype)
        self.W = a["weights"].astype(dtype)
        self.step = 0
        cols = [(col0, size, np.asarray(tables[node], dtype)) for node, col0, size in plan.tables]
        n_steps = min(t.shape[0] for _, _, t in cols) if cols else 0
        self.tables = np.zeros((n_steps, self.nt), dtype)
        for col0, size, t in cols:
            self.tables[:, col0:col0 + size] = t[:n_steps]
        for ens, (row0, n) in plan.ens_state.items():
            self.v[row0:row0 + n] = model.initial_voltage(ens, trial_seed)
        for ens, (row0, n, dims) in plan.learned_enc.items():
            self.lenc[row0:row0 + n * dims] = model.params[ens].scaled_encoders.reshape(-1)
        for conn, (row0, so, n) in plan.learned_dec.items():
            self.ldec[row0:row0 + so * n] = np.asarray(model.params[conn].weights).reshape(-1)
        s64 = a["cleanup_s64"]
        self.grids, off = [], 0
        for d in a["cleanup"]:
            G, dims = int(d[0]), int(d[1])
            self.grids.append(s64[off:off + G * dims].reshape(G, dims))
            off += G * dims
        self.cidx = np.zeros(len(self.grids), np.int64)
        self.probe_rows = []
        self.ntypes = a["ntypes"].astype(np.float64)
        self._build_row_programs()
        self._build_small()

    # ------------------------------------------------------------------ merged row programs
    def _csr(self, srcs, parity):
        """Rows ``srcs`` of the plan's CSR program as ONE scipy matrix over the vec rows of steps with this parity."""
        ptr, ent = self.p.arrays["csr_ptr"], self.p.arrays["csr_ent%d" % parity]
        val = np.ascontiguousarray(ent[:, 1]).view(np.float32).astype(self.dt_)
        indptr, idx, data = [0], [], []
        for s in srcs:
            lo, hi = int(ptr[s]), int(ptr[s + 1])
            idx.append(ent[lo:hi, 0])
            data.append(val[lo:hi])
            indptr.append(indptr[-1] + hi - lo)
        idx = np.concatenate(idx) if idx else np.zeros(0, np.int64)
        data = np.concatenate(data) if data else np.zeros(0, self.dt_)
        return sp.csr_matrix((data, idx, np.asarray(indptr)), shape=(len(srcs), self.nv))

    def _build_row_programs(self):
        a = self.p.arrays
        rows = a["lin_rows"]
        self.level_prog = []
        for st in a["stages"]:
            seg = rows[st[10]:st[10] + st[11]]
            cur, prev = seg[seg[:, 1] == 3], seg[seg[:, 1] == 4]
            # kind 3 rows read what this step reads; kind 4 rows the half the previous step read (the other parity's map)
            self.level_prog.append([(self._csr(cur[:, 0], par), cur[:, 2].astype(np.int64),
                                     self._csr(prev[:, 0], 1 - par), prev[:, 2].astype(np.int64)) for par in (0, 1)])
        lin0, n_lin = int(self.p.scalars["lin0"]), int(self.p.scalars["n_lin"])
        fin, ab = rows[lin0:lin0 + n_lin], a["lin_ab"][lin0:lin0 + n_lin].astype(self.dt_)
        k0, k1, k2, k5 = (fin[:, 1] == k for k in (0, 1, 2, 5))
        self.f_dst, self.f_a, self.f_b = fin[k0][:, 2].astype(np.int64), ab[k0][:, 0], ab[k0][:, 1]
        self.f_mat = [self._csr(fin[k0][:, 0], par) for par in (0, 1)]
        self.p_dst = fin[k1][:, 2].astype(np.int64)
        self.p_mat = [self._csr(fin[k1][:, 0], par) for par in (0, 1)]
        self.t_src, self.t_dst = fin[k2][:, 0].astype(np.int64), fin[k2][:, 2].astype(np.int64)
        self.t_a, self.t_b = ab[k2][:, 0], ab[k2][:, 1]
        self.n5_src, self.n5_dst = fin[k5][:, 0].astype(np.int64), fin[k5][:, 2].astype(np.int64)

    def _build_small(self):
        """All narrow ensembles of a level as one concatenated population."""
        a, W = self.p.arrays, self.W
        self.small = []
        for st in a["stages"]:
            desc = a["ens_small"][st[0]:st[0] + st[1]]
            if len(desc) == 0:
                self.small.append(None)
                continue
            N = int(desc[:, 0].sum())
            bias = np.zeros(N, self.dt_)
            sidx, tid = np.zeros(N, np.int64), np.zeros(N, np.int64)
            e_r, e_c, e_v, d_r, d_c, d_v, out_rows = [], [], [], [], [], [], []
            o = 0
            for d in desc:
                n, dims, nout, s0, w_off, in_row0, out_vec, t, stride = (int(x) for x in d)
                pk = W[w_off:w_off + n * stride].reshape(n, stride)
                bias[o:o + n] = pk[:, 0]
                nidx = o + np.arange(n)
                for k in range(dims):                       # encoders: neuron row <- the ensemble's input rows
                    e_r.append(nidx)
                    e_c.append(np.full(n, in_row0 + k))
                    e_v.append(pk[:, 1 + k])
                for j in range(nout):                       # decoders: output slot <- the ensemble's neurons
                    d_r.append(np.full(n, len(out_rows)))
                    d_c.append(nidx)
                    d_v.append(pk[:, 1 + dims + j])
                    out_rows.append(out_vec + j)
                sidx[o:o + n] = s0 + np.arange(n)
                tid[o:o + n] = t
                o += n
            # what nengo's operator merging builds for the per-ensemble DotIncs: one block-sparse matrix each way
            enc = sp.csr_matrix((np.concatenate(e_v), (np.concatenate(e_r), np.concatenate(e_c))), shape=(N, self.nv))
            dec = sp.csr_matrix((np.concatenate(d_v), (np.concatenate(d_r), np.concatenate(d_c))), shape=(len(out_rows), N))
            tids = np.unique(tid)
            self.small.append(dict(bias=bias, enc=enc, dec=dec, sidx=sidx, tid=int(tids[0]) if len(tids) == 1 else tid,
                                   out_rows=np.asarray(out_rows, np.int64)))

    # ------------------------------------------------------------------ neurons (nengo/neurons.py step functions)
    def _neurons(self, tid, J, v, r):
        """Vectorised over a population whose members may have different neuron types (``tid`` per neuron or scalar)."""
        out = np.zeros_like(J)
        tids = np.unique(tid) if isinstance(tid, np.ndarray) else [tid]
        for t in tids:
            m = slice(None) if len(tids) == 1 else (tid == t)
            kind, tau_rc, tau_ref, min_v, amp = self.ntypes[int(t)][:5]
            Jm = J[m]
            if int(kind) == NT_LIF:
                vm, rm = v[m], r[m]
                rm = rm - self.dt
                delta = np.clip(self.dt - rm, 0, self.dt)
                vm = vm - (Jm - vm) * np.expm1(-delta / tau_rc)
                spiked = vm > 1
                o = spiked * (amp / self.dt)
                t_spike = self.dt + tau_rc * np.log1p(-(vm[spiked] - 1) / (Jm[spiked] - 1))
                vm[vm < min_v] = min_v
                vm[spiked] = 0
                rm[spiked] = tau_ref + t_spike
                v[m], r[m] = vm, rm
                out[m] = o
            elif int(kind) == NT_LIFRATE:
                j = Jm - 1
                o = np.zeros_like(Jm)
                pos = j > 0
                o[pos] = amp / (tau_ref + tau_rc * np.log1p(1.0 / j[pos]))
                out[m] = o
            else:
                out[m] = amp * np.maximum(Jm, 0)
        return out

    # ------------------------------------------------------------------ stepping
    def run_steps(self, n):
        for _ in range(n):
            self.one_step()

    def one_step(self):
        a, W, s = self.p.arrays, self.W, self.step
        par = s & 1
        par_old, par_new = (self.nf, 0) if par else (0, self.nf)
        if self.nt:
            r0 = self.tab_row0 + (self.nt if par else 0)
            self.vec[r0:r0 + self.nt] = self.tables[s]
        for lvl, st in enumerate(a["stages"]):
            m_cur, d_cur, m_prev, d_prev = self.level_prog[lvl][par]
            if len(d_cur):
                self.vec[d_cur] = m_cur @ self.vec
            if len(d_prev):
                self.vec[d_prev] = m_prev @ self.vec
            sm = self.small[lvl]
            if sm is not None:
                J = sm["bias"] + sm["enc"] @ self.vec
                v, r = self.v[sm["sidx"]], self.ref[sm["sidx"]]
                out = self._neurons(sm["tid"], J, v, r)
                self.v[sm["sidx"]], self.ref[sm["sidx"]] = v, r
                self.vec[sm["out_rows"]] = sm["dec"] @ out
            for d in a["ens_big"][st[2]:st[2] + st[3]]:
                (n, dims, dpad, s0, act0, enc_off, bias_off, in_row0, tid, flags, jn_row0, jn_m, jn_w, voja_row,
                 scale_off, alpha_bits) = (int(x) for x in d)
                x = self.vec[in_row0:in_row0 + dims].copy()
                if flags & 1:
                    E = self.lenc[enc_off:enc_off + n * dims].reshape(n, dims)
                else:
                    E = W[enc_off:enc_off + n * dpad].reshape(n, dpad)[:, :dims]
                J = W[bias_off:bias_off + n] + E @ x
                if jn_m:
                    J = J + W[jn_w:jn_w + n * jn_m].reshape(n, jn_m) @ self.vec[jn_row0:jn_row0 + jn_m]
                out = self._neurons(tid, J, self.v[s0:s0 + n], self.ref[s0:s0 + n])
                self.act[act0:act0 + n] = out
                if flags & 1:
                    aL = float(np.int32(alpha_bits).view(np.float32)) * self.vec[voja_row]
                    if aL != 0.0:
                        fired = np.flatnonzero(out)
                        if fired.size:
                            sc = W[scale_off:scale_off + n][fired]
                            o = out[fired]
                            E[fired] += aL * (sc[:, None] * np.outer(o, x) - o[:, None] * E[fired])    # E is a view of lenc
            for ci in range(st[6], st[6] + st[7]):
                G, dims, dpad, s_off, in_row0, out_vec = (int(x) for x in a["cleanup"][ci])
                g = int(np.argmax(self.grids[ci] @ self.vec[in_row0:in_row0 + dims].astype(np.float64)))
                self.cidx[ci] = g
                self.vec[out_vec:out_vec + dims] = W[s_off + g * dpad:s_off + g * dpad + dims]
            for d in a["gate"][st[8]:st[8] + st[9]]:
                dims, in_row0, out_vec = int(d[0]), int(d[1]), int(d[2])
                rate, thres, atol = (float(np.int32(x).view(np.float32)) for x in d[3:6])
                x = self.vec[in_row0:in_row0 + 2 * dims + 1]
                p_, q_ = x[:dims], x[dims:2 * dims]
                open_ = abs(x[-1]) <= atol and float(p_ @ q_) > thres
                self.vec[out_vec:out_vec + dims] = rate * (p_ - q_) if open_ else 0.0
            for d in a["dec"][st[4]:st[4] + st[5]]:
                n, so, jpad, act0, w_off, out_vec = (int(x) for x in d[:6])
                act = self.act[act0:act0 + n]
                nz = np.flatnonzero(act)                                  # spikes are sparse
                Wd = W[w_off:w_off + n * jpad].reshape(n, jpad)
                self.vec[out_vec:out_vec + so] = act[nz] @ Wd[nz, :so]
        for d in a["pes"]:
            n, so, d_off, a_off, act0, err_row0, out_vec = (int(x) for x in d[:7])
            alpha = float(np.int32(d[7]).view(np.float32))
            D = self.ldec[d_off:d_off + so * n].reshape(so, n)
            if s > 0:
                D += np.outer(alpha * self.vec[err_row0:err_row0 + so], self.afilt[1 - par, a_off:a_off + n])
            act = self.act[act0:act0 + n]
            nz = np.flatnonzero(act)
            self.vec[out_vec:out_vec + so] = D[:, nz] @ act[nz]
        probe = np.zeros(int(self.p.scalars["n_probe"]), self.dt_)
        new_f = self.f_b * (self.f_mat[par] @ self.vec) + self.f_a * self.vec[1 + self.f_dst + par_old]
        if len(self.p_dst):
            probe[self.p_dst] = self.p_mat[par] @ self.vec
        if len(self.n5_dst):
            probe[self.n5_dst] = self.act[self.n5_src]
        if len(self.t_dst):
            self.afilt[1 - par, self.t_dst] = self.t_b * self.act[self.t_src] + self.t_a * self.afilt[par, self.t_dst]
        self.vec[1 + self.f_dst + par_new] = new_f
        self.probe_rows.append(probe)
        self.step += 1

    def probe_data(self, info):
        rows = np.stack(self.probe_rows)[:, info.row0:info.row0 + info.size]
        return rows[info.period - 1::info.period] if info.period > 1 else rows
