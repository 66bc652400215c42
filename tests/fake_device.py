"""TEST-ONLY stand-in for ``membrane_solver_b200.context.DeviceMesh`` backed by the host emulator
(tests/emul): lets the CPU tier drive the plugin modules, the evaluation manager and the
residency logic end to end without a GPU.  The product never imports this."""

import numpy as np

import ms_test_helpers as H
from membrane_solver_b200 import _lib as L
from membrane_solver_b200.context import EvalResult


class FakeDeviceMesh:
    instances = []

    def __init__(self, device=0, **pack):
        self.pack = {k: v for k, v in pack.items() if v is not None and k in ("threads", "max_owned", "max_local")}
        self.nv = self.nf = 0
        self.gamma = 1.0
        self.kappa = self.c0 = 0.0
        self.k_tilt = 0.0
        self.tilts = None
        self.arrays = {}
        self.topology_uploads = 0
        self.evals = 0
        FakeDeviceMesh.instances.append(self)

    # -- topology / parameters
    def set_topology(self, nv, tri, *, is_boundary=None, body_mask=None, fixed_mask=None, n_owned=None,
                     order_hint=None):
        self.nv, self.tri = int(nv), np.ascontiguousarray(tri, dtype=np.int32).reshape(-1, 3)
        self.nf = self.tri.shape[0]
        self.is_boundary, self.body_mask, self.fixed = is_boundary, body_mask, fixed_mask
        self.topology_uploads += 1

    def set_surface_tension(self, gamma):
        self.gamma = gamma

    def set_bending_params(self, kappa, c0):
        self.kappa, self.c0 = kappa, c0

    def set_tilt_rigidity(self, k):
        self.k_tilt = float(k)

    def set_tilts(self, t):
        self.tilts = np.array(t, dtype=np.float64)

    options = staticmethod(lambda modules, **kw: dict(modules=modules, **kw))

    # -- evaluation
    def eval_host(self, opts, pos, *, grad=None, volgrad=None, tilt_grad=None):
        self.evals += 1
        kw = dict(self.pack)
        g, k, c = self.gamma, self.kappa, self.c0
        out = H.emulate(pos, self.tri, modules=opts["modules"], flags=opts.get("flags", 0),
                        want_grad=opts.get("want_grad", True), is_boundary=self.is_boundary, body_mask=self.body_mask,
                        tilts=self.tilts, gamma=g if np.ndim(g) else None, gamma_u=g if not np.ndim(g) else 1.0,
                        kappa=k if np.ndim(k) else None, kappa_u=0.0 if np.ndim(k) else float(k),
                        c0=c if np.ndim(c) else None, c0_u=0.0 if np.ndim(c) else float(c), k_tilt=self.k_tilt, **kw)
        sc = np.zeros(L.SC_COUNT)
        sc[L.SC_E_SURFACE], sc[L.SC_AREA], sc[L.SC_VOLUME] = out["E_surface"], out["area"], out["volume"]
        sc[L.SC_E_BENDING], sc[L.SC_E_TILT] = out["E_bending"], out["E_tilt"]
        sc[L.SC_E_BENDING_TILT] = out["E_bending_tilt"]
        g_out, gc = out["grad"], out["volgrad"]
        if opts.get("want_grad", True):
            # numpy restatement of k_dots / k_project (TEST ONLY)
            sc[L.SC_G_G], sc[L.SC_G_GC], sc[L.SC_GC_GC] = (g_out * g_out).sum(), (g_out * gc).sum(), (gc * gc).sum()
            mode = opts.get("constraint_mode", -1)
            if mode == 0 and (opts["modules"] & L.MOD_VOLUME) and sc[L.SC_GC_GC] > 1e-18:
                lam = sc[L.SC_G_GC] / sc[L.SC_GC_GC]
                g_out = g_out - lam * gc
                sc[L.SC_LAMBDA] = lam
            elif mode == 1 and (opts["modules"] & L.MOD_VOLUME):
                coef = opts.get("k_vol", 0.0) * (sc[L.SC_VOLUME] - opts.get("v_target", 0.0))
                g_out = g_out + coef * gc
                sc[L.SC_LAMBDA] = coef
            if opts.get("apply_fixed") and self.fixed is not None:
                g_out = np.where(np.asarray(self.fixed, bool)[:, None], 0.0, g_out)
        self.arrays = {L.ARR_GRAD: g_out, L.ARR_VOLGRAD: gc, L.ARR_TILT_GRAD: out["tilt_grad"],
                       L.ARR_E_VERTEX: out["e_vertex"], L.ARR_K_VECS: out["k_vecs"], L.ARR_A_VOR: out["a_vor"],
                       L.ARR_A_EFF: out["a_eff"]}
        if grad is not None:
            grad[:] = g_out
        if volgrad is not None:
            volgrad[:] = gc
        if tilt_grad is not None:
            tilt_grad[:] = out["tilt_grad"]
        return EvalResult(sc)

    def download(self, which):
        return np.array(self.arrays[which])

    def close(self):
        pass
