"""TEST INFRASTRUCTURE: NumPy interpreter of a lowered ``DevicePlan`` for ONE trial.

It executes the same launch sequence and descriptor semantics as the CUDA kernels
(``csrc/ssb_kernels.cuh``) so that the host-side lowering can be checked against the
operator-level oracle on a machine without a GPU.  It is never imported by the package."""
import numpy as np

from sspslam_b200.lowering import NT_LIF, NT_LIFRATE


class PlanInterpreter:
    def __init__(self, plan, model, network, tables, dtype=np.float64, trial_seed=None):
        self.p, self.dt_ = plan, dtype
        a, sc = plan.arrays, plan.scalars
        self.dt = dtype(sc["dt"])
        self.nf, self.nt, self.tab_row0 = int(sc["nf"]), int(sc["nt"]), int(sc["tab_row0"])
        self.vec = np.zeros(int(sc["nv"]), dtype)
        self.vec[0] = 1
        self.v = np.zeros(int(sc["nn"]), dtype)
        self.ref = np.zeros(int(sc["nn"]), dtype)
        self.act = np.zeros(max(1, int(sc["n_act"])), dtype)
        self.lenc = np.zeros(max(1, int(sc["n_lenc"])), dtype)
        self.ldec = np.zeros(max(1, int(sc["n_ldec"])), dtype)
        self.afilt = np.zeros((2, max(1, int(sc["n_afilt"]))), dtype)
        self.W = a["weights"].astype(dtype)
        # plans lowered with per_trial=True: this trial's seed-dependent weights (its own model) in their arenas
        self.per_trial = bool(sc.get("per_trial_weights", 0))
        if self.per_trial:
            from sspslam_b200 import lowering
            wp, enc, dec = lowering.trial_weights(network, model)
            assert wp.size == a["weights_pt"].size
            self.Wp = wp.astype(dtype)
            for ens, (row0, n, dims) in plan.pt_enc.items():
                self.lenc[row0:row0 + n * dims] = enc[ens].reshape(-1)
            for conn, (row0, so, n) in plan.pt_dec.items():
                self.ldec[row0:row0 + so * n] = dec[conn].reshape(-1)
        else:
            self.Wp = self.W
        self.ptr = a["csr_ptr"]
        self.ent = [a["csr_ent0"], a["csr_ent1"]]          # (vec row, coefficient bits), per step parity
        self.val = np.ascontiguousarray(a["csr_ent0"][:, 1]).view(np.float32).astype(dtype)
        assert np.array_equal(a["csr_ent0"][:, 1], a["csr_ent1"][:, 1])
        self.step = 0
        self.tables = np.zeros((0, self.nt), dtype)
        cols = []
        for node, col0, size in plan.tables:
            cols.append((col0, size, np.asarray(tables[node], dtype)))
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

    # ---- CSR rows: ``which`` = 0 reads what this step reads, 1 the half the previous step read
    def rows(self, row0, n, which=0):
        ent = self.ent[(self.step & 1) ^ which]
        out = np.zeros(n, self.dt_)
        for r in range(n):
            lo, hi = self.ptr[row0 + r], self.ptr[row0 + r + 1]
            acc = self.dt_(0)
            for k in range(lo, hi):
                acc += self.val[k] * self.vec[int(ent[k, 0])]
            out[r] = acc
        return out

    def neuron(self, tid, J, s0, n):
        kind, tau_rc, tau_ref, min_v, amp = self.p.arrays["ntypes"][tid].astype(np.float64)[:5]
        dt = self.dt
        if int(kind) == NT_LIF:
            v, r = self.v[s0:s0 + n], self.ref[s0:s0 + n]
            r -= dt
            delta = np.clip(dt - r, 0, dt)
            v -= (J - v) * np.expm1(-delta / tau_rc)
            spiked = v > 1
            out = spiked * (amp / dt)
            t_spike = dt + tau_rc * np.log1p(-(v[spiked] - 1) / (J[spiked] - 1))
            v[v < min_v] = min_v
            v[spiked] = 0
            r[spiked] = tau_ref + t_spike
            return out.astype(self.dt_)
        if int(kind) == NT_LIFRATE:
            j = J - 1
            out = np.zeros_like(J)
            pos = j > 0
            out[pos] = amp / (tau_ref + tau_rc * np.log1p(1.0 / j[pos]))
            return out
        return amp * np.maximum(J, 0)

    def run_steps(self, n):
        for _ in range(n):
            self.one_step()

    def one_step(self):
        a, W = self.p.arrays, self.W
        s = self.step
        par_old, par_new = (self.nf, 0) if s & 1 else (0, self.nf)
        if self.nt:
            r0 = self.tab_row0 + (self.nt if s & 1 else 0)                       # k_begin: the copy of this step's parity
            self.vec[r0:r0 + self.nt] = self.tables[s]
        for st in a["stages"]:
            for src, kind, dst in a["lin_rows"][st[10]:st[10] + st[11]]:     # materialise this level's sink rows
                assert kind in (3, 4)
                self.vec[dst] = self.rows(src, 1, 1 if kind == 4 else 0)[0]
                if getattr(self, "fused_next", None) is not None and kind == 3 and int(dst) in self.fused_next:
                    want = self.vec[dst]                                     # what the previous step's fused row predicted
                    got = self.fused_next.pop(int(dst))
                    self.fused_err = max(getattr(self, "fused_err", 0.0), abs(got - want) / (abs(want) + 1e-3))
                    self.fused_checked = getattr(self, "fused_checked", 0) + 1
            for d in a["ens_small"][st[0]:st[0] + st[1]]:
                n, dims, nout, s0, w_off, in_row0, out_vec, tid, stride = (int(x) for x in d)
                pk = self.Wp[w_off:w_off + n * stride].reshape(n, stride)
                x = self.vec[in_row0:in_row0 + dims].copy()
                out = self.neuron(tid, pk[:, 0] + pk[:, 1:1 + dims] @ x, s0, n)
                self.vec[out_vec:out_vec + nout] = pk[:, 1 + dims:1 + dims + nout].T @ out
            for d in a["ens_big"][st[2]:st[2] + st[3]]:
                (n, dims, dpad, s0, act0, enc_off, bias_off, in_row0, tid, flags, jn_row0, jn_m, jn_w, voja_row,
                 scale_off, alpha_bits) = (int(x) for x in d)
                x = self.vec[in_row0:in_row0 + dims].copy()
                if flags & 1:
                    E = self.lenc[enc_off:enc_off + n * dims].reshape(n, dims)
                else:
                    E = W[enc_off:enc_off + n * dpad].reshape(n, dpad)[:, :dims]
                Wb = self.Wp if flags & 4 else W       # bit 2: bias / Voja scale / neuron-current weights are per trial
                J = Wb[bias_off:bias_off + n] + E @ x
                if jn_m:
                    J = J + Wb[jn_w:jn_w + n * jn_m].reshape(n, jn_m) @ self.vec[jn_row0:jn_row0 + jn_m]
                out = self.neuron(tid, J, s0, n)
                self.act[act0:act0 + n] = out
                if flags & 1:
                    aL = np.int32(alpha_bits).view(np.float32).astype(self.dt_) * self.vec[voja_row]
                    sc = Wb[scale_off:scale_off + n]
                    E += aL * (sc[:, None] * np.outer(out, x) - out[:, None] * E)   # E is a view of lenc
            for ci in range(st[6], st[6] + st[7]):
                G, dims, dpad, s_off, in_row0, out_vec = (int(x) for x in a["cleanup"][ci])
                x = self.vec[in_row0:in_row0 + dims].copy()
                g = int(np.argmax(self.grids[ci] @ x.astype(np.float64)))
                self.cidx[ci] = g
                self.vec[out_vec:out_vec + dims] = W[s_off + g * dpad:s_off + g * dpad + dims]
            for d in a["gate"][st[8]:st[8] + st[9]]:
                dims, in_row0, out_vec = int(d[0]), int(d[1]), int(d[2])
                rate, thres, atol = (np.int32(x).view(np.float32).astype(np.float64) for x in d[3:6])
                x = self.vec[in_row0:in_row0 + 2 * dims + 1].copy()
                p_, q_ = x[:dims], x[dims:2 * dims]
                open_ = abs(x[-1]) <= atol and float(p_ @ q_) > thres
                self.vec[out_vec:out_vec + dims] = rate * (p_ - q_) if open_ else 0.0
            for d in a["dec"][st[4]:st[4] + st[5]]:
                n, so, jpad, act0, w_off, out_vec, nch = (int(x) for x in d[:7])
                if self.per_trial:                         # per-trial static decoder: ldec rows (k_decode_pt)
                    Wd = self.ldec[w_off:w_off + so * n].reshape(so, n).T
                else:
                    Wd = W[w_off:w_off + n * jpad].reshape(n, jpad)[:, :so]
                per = -(-n // nch)
                total = np.zeros(so, self.dt_)
                for c in range(nch):                       # split-K partial sums, added in chunk order
                    lo, hi = c * per, min(n, (c + 1) * per)
                    total = total + Wd[lo:hi].T @ self.act[act0 + lo:act0 + hi]
                self.vec[out_vec:out_vec + so] = total
        for d in a["pes"]:
            n, so, d_off, a_off, act0, err_row0, out_vec = (int(x) for x in d[:7])
            alpha = np.int32(d[7]).view(np.float32).astype(self.dt_)
            nch = int(d[10])
            D = self.ldec[d_off:d_off + so * n].reshape(so, n)
            if s > 0:
                err = self.vec[err_row0:err_row0 + so].copy()
                D += np.outer(alpha * err, self.afilt[1 - (s & 1), a_off:a_off + n])
            per = -(-n // nch)
            total = np.zeros(so, self.dt_)
            for c in range(nch):
                lo, hi = c * per, min(n, (c + 1) * per)
                total = total + D[:, lo:hi] @ self.act[act0 + lo:act0 + hi]
            self.vec[out_vec:out_vec + so] = total
        probe = np.zeros(int(self.p.scalars["n_probe"]), self.dt_)
        new_f = {}
        lin0 = int(self.p.scalars["lin0"])
        n_lin = int(self.p.scalars["n_lin"])
        for (src, kind, dst), (ca, cb) in zip(a["lin_rows"][lin0:lin0 + n_lin], a["lin_ab"][lin0:lin0 + n_lin].astype(self.dt_)):
            if kind == 0:
                new_f[1 + dst + par_new] = cb * self.rows(src, 1)[0] + ca * self.vec[1 + dst + par_old]
            elif kind == 1:
                probe[dst] = self.rows(src, 1)[0]
            elif kind == 5:
                probe[dst] = self.act[src]
            else:
                ob = s & 1
                self.afilt[1 - ob, dst] = cb * self.act[src] + ca * self.afilt[ob, dst]
        # fused end-of-step rows (lowering: step fusion): the NEXT step's level-0 sink rows evaluated now, from this step's
        # columns and the next step's table rows (the other parity's copy, as the device's input prefetch leaves them)
        n_fused = int(self.p.scalars.get("n_lin_fused", 0))
        self.fused_next = None
        if n_fused and self.nt and s + 1 < len(self.tables):
            r1 = self.tab_row0 + (0 if s & 1 else self.nt)
            self.vec[r1:r1 + self.nt] = self.tables[s + 1]
            n_all = len(a["lin_rows"])
            self.fused_next = {int(dst): self.rows(int(src), 1)[0] for src, kind, dst in a["lin_rows"][n_all - n_fused:]}
        for k, val in new_f.items():
            self.vec[k] = val
        self.probe_rows.append(probe)
        self.step += 1

    def probe_data(self, info):
        rows = np.stack(self.probe_rows)[:, info.row0:info.row0 + info.size]
        return rows[info.period - 1::info.period] if info.period > 1 else rows
