"""ctypes binding of libsimplemcmc_cuda.so (include/simplemcmc.h).  This is the same call sequence the Julia
shim performs with `ccall`.  There is no CPU fallback: if the library or a CUDA device is
missing every entry point raises."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsimplemcmc_cuda.so")

c_double_p = ctypes.POINTER(ctypes.c_double)


class SMCError(RuntimeError):
    def __init__(self, status, message):
        RuntimeError.__init__(self, message)
        self.status = status


class SamplerConfig(ctypes.Structure):
    _fields_ = [("steps", ctypes.c_int64), ("burnin", ctypes.c_int64), ("isteps", ctypes.c_int64),
                ("stepsize", ctypes.c_double), ("seed", ctypes.c_uint64), ("chain_offset", ctypes.c_int64),
                ("store_draws", ctypes.c_int32), ("init_per_chain", ctypes.c_int32),
                ("inject_normals", c_double_p), ("inject_uniforms", c_double_p),
                ("inject_uniforms_per_step", ctypes.c_int64)]


class RunOutput(ctypes.Structure):
    _fields_ = [("draws", c_double_p), ("loglik", c_double_p), ("accept", c_double_p), ("final_state", c_double_p),
                ("accept_rate", c_double_p), ("mean", c_double_p), ("m2", c_double_p), ("epsilon", c_double_p),
                ("jmax", c_double_p), ("device_ms", ctypes.c_double), ("n_evals", ctypes.c_int64),
                ("n_leapfrogs", ctypes.c_int64), ("n_kernel_launches", ctypes.c_int64),
                ("first", c_double_p), ("lag1", c_double_p)]


EXPORTS = ["smc_abi_version", "smc_last_error", "smc_ctx_create", "smc_ctx_destroy", "smc_ctx_comm_init",
           "smc_comm_unique_id", "smc_ctx_allreduce", "smc_ctx_p2p_export", "smc_ctx_p2p_import", "smc_model_compile", "smc_model_destroy", "smc_model_bind_data",
           "smc_model_finalize", "smc_model_nparams", "smc_eval", "smc_eval_timed", "smc_hmc_run", "smc_rwm_run",
           "smc_nuts_run", "smc_philox_normals", "smc_philox_uniforms", "smc_vm_plan_describe"]

_lib = None


def load():
    """dlopen the library (building is __graft_entry__.build()'s / build.py's job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SMCError(-2, "libsimplemcmc_cuda.so is not built (%s): run `python simplemcmc.jl_b200/build.py`; "
                           "there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    lib.smc_last_error.restype = ctypes.c_char_p
    vp, i64, dbl, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int
    lib.smc_vm_plan_describe.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    lib.smc_ctx_create.argtypes = [i32, ctypes.POINTER(vp)]
    lib.smc_ctx_destroy.argtypes = [vp]
    lib.smc_ctx_comm_init.argtypes = [vp, ctypes.c_char_p, i32, i32, i32]
    lib.smc_comm_unique_id.argtypes = [ctypes.c_char_p, i32]
    lib.smc_ctx_allreduce.argtypes = [vp, c_double_p, i64]
    lib.smc_model_compile.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.smc_model_destroy.argtypes = [vp]
    lib.smc_model_bind_data.argtypes = [vp, ctypes.c_char_p, c_double_p, ctypes.POINTER(i64), i32]
    lib.smc_model_finalize.argtypes = [vp, i64, i32]
    lib.smc_model_nparams.argtypes = [vp, ctypes.POINTER(i64)]
    lib.smc_eval.argtypes = [vp, c_double_p, i64, c_double_p, c_double_p]
    lib.smc_eval_timed.argtypes = [vp, i32, i32, c_double_p, c_double_p]
    for fn in ("smc_hmc_run", "smc_rwm_run", "smc_nuts_run"):
        getattr(lib, fn).argtypes = [vp, ctypes.POINTER(SamplerConfig), c_double_p, i64, ctypes.POINTER(RunOutput)]
    lib.smc_philox_normals.argtypes = [ctypes.c_uint64, i64, i64, ctypes.c_int32, i64, c_double_p]
    lib.smc_philox_uniforms.argtypes = [ctypes.c_uint64, i64, i64, ctypes.c_int32, i64, c_double_p]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SMCError(rc, load().smc_last_error().decode("utf-8", "replace"))


def dptr(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


class Context(object):
    _default = {}

    def __init__(self, device=0):
        self.lib = load()
        self.h = ctypes.c_void_p()
        check(self.lib.smc_ctx_create(int(device), ctypes.byref(self.h)))
        self.device = device
        self.rank, self.nranks = 0, 1

    @classmethod
    def default(cls, device=0):
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def comm_init(self, unique_id, rank, nranks):
        check(self.lib.smc_ctx_comm_init(self.h, unique_id, len(unique_id), rank, nranks))
        self.rank, self.nranks = rank, nranks

    def p2p_export(self, nranks):
        """64-byte IPC handle of this rank's peer-exchange buffer (row-sharded models on one node, smc_p2p.cu)"""
        buf = ctypes.create_string_buffer(64)
        check(self.lib.smc_ctx_p2p_export(self.h, int(nranks), buf, 64))
        return buf.raw

    def p2p_import(self, handles, nranks, rank):
        """handles: the nranks exported handles concatenated in rank order"""
        assert len(handles) == 64 * nranks
        check(self.lib.smc_ctx_p2p_import(self.h, handles, int(nranks), int(rank)))

    def allreduce(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        check(self.lib.smc_ctx_allreduce(self.h, dptr(arr), arr.size))
        return arr

    def close(self):
        if self.h:
            self.lib.smc_ctx_destroy(self.h)
            self.h = None


def unique_id():
    buf = ctypes.create_string_buffer(128)
    check(load().smc_comm_unique_id(buf, 128))
    return buf.raw


class DeviceModel(object):
    """a compiled tape with its data bound on the device"""

    def __init__(self, ctx, tape, bound, max_chains=1, row_shard=False):
        self.ctx, self.lib = ctx, ctx.lib
        self.h = ctypes.c_void_p()
        check(self.lib.smc_model_compile(ctx.h, tape, len(tape), ctypes.byref(self.h)))
        self.h2d_bytes = 0
        for name, arr in bound:
            self.bind(name, arr)
        self.max_chains = 0
        self.row_shard = row_shard
        self.finalize(max_chains)
        n = ctypes.c_int64()
        check(self.lib.smc_model_nparams(self.h, ctypes.byref(n)))
        self.nparams = n.value

    def bind(self, name, arr):
        a = np.asfortranarray(arr, dtype=np.float64)
        dims = (ctypes.c_int64 * max(a.ndim, 1))(*a.shape)
        check(self.lib.smc_model_bind_data(self.h, name.encode(), dptr(a), dims, a.ndim))
        self.h2d_bytes += a.nbytes

    def finalize(self, max_chains):
        check(self.lib.smc_model_finalize(self.h, int(max_chains), 1 if self.row_shard else 0))
        self.max_chains = int(max_chains)

    def ensure_chains(self, C):
        if C > self.max_chains:
            self.finalize(C)

    def eval(self, beta, gradient=True):
        beta = np.ascontiguousarray(np.atleast_2d(beta), dtype=np.float64)
        C, M = beta.shape
        if M != self.nparams:
            raise SMCError(-1, "parameter vector has %d elements, the model has %d" % (M, self.nparams))
        self.ensure_chains(C)
        logp = np.empty(C)
        grad = np.empty((C, M)) if gradient else None
        check(self.lib.smc_eval(self.h, dptr(beta), C, dptr(logp), dptr(grad)))
        return logp, grad

    def eval_timed(self, reps=10, flush_l2=True):
        a, b = ctypes.c_double(), ctypes.c_double()
        check(self.lib.smc_eval_timed(self.h, reps, 1 if flush_l2 else 0, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def run(self, sampler, init, chains, steps, burnin, isteps=1, stepsize=0.0, seed=0, chain_offset=0,
            store_draws=True, want=("draws", "loglik", "accept"), inject_normals=None, inject_uniforms=None):
        init = np.ascontiguousarray(init, dtype=np.float64)
        per_chain = init.ndim == 2
        C, M = int(chains), self.nparams
        if per_chain and init.shape != (C, M):
            raise SMCError(-1, "init must be [M] or [chains][M]")
        self.ensure_chains(C)
        samples = steps - burnin
        cfg = SamplerConfig(steps=steps, burnin=burnin, isteps=isteps, stepsize=stepsize, seed=seed,
                            chain_offset=chain_offset, store_draws=1 if store_draws else 0,
                            init_per_chain=1 if per_chain else 0)
        keep = []
        if inject_normals is not None:
            n = np.ascontiguousarray(inject_normals, dtype=np.float64)
            assert n.shape == (steps, C, M), n.shape
            cfg.inject_normals = dptr(n)
            keep.append(n)
        if inject_uniforms is not None:
            u = np.ascontiguousarray(inject_uniforms, dtype=np.float64)
            assert u.ndim == 3 and u.shape[:2] == (steps, C), u.shape
            cfg.inject_uniforms = dptr(u)
            cfg.inject_uniforms_per_step = u.shape[2]
            keep.append(u)
        res = {}
        out = RunOutput()
        samples = max(samples, 0)
        if store_draws and "draws" in want:
            res["draws"] = np.empty((samples, C, M))
            out.draws = dptr(res["draws"])
        if "loglik" in want:
            res["loglik"] = np.empty((samples, C))
            out.loglik = dptr(res["loglik"])
        if "accept" in want:
            res["accept"] = np.empty((samples, C))
            out.accept = dptr(res["accept"])
        res["final_state"] = np.empty((C, M))
        res["accept_rate"] = np.empty(C)
        res["mean"] = np.empty((C, M))
        res["m2"] = np.empty((C, M))
        res["first"] = np.empty((C, M))
        res["lag1"] = np.empty((C, M))
        out.final_state, out.accept_rate = dptr(res["final_state"]), dptr(res["accept_rate"])
        out.mean, out.m2 = dptr(res["mean"]), dptr(res["m2"])
        out.first, out.lag1 = dptr(res["first"]), dptr(res["lag1"])
        if sampler == "nuts":
            res["epsilon"] = np.empty((steps, C))
            res["jmax"] = np.empty((steps, C))
            out.epsilon, out.jmax = dptr(res["epsilon"]), dptr(res["jmax"])
        fn = {"hmc": self.lib.smc_hmc_run, "rwm": self.lib.smc_rwm_run, "nuts": self.lib.smc_nuts_run}[sampler]
        check(fn(self.h, ctypes.byref(cfg), dptr(init), C, ctypes.byref(out)))
        res.update(device_ms=out.device_ms, n_evals=out.n_evals, n_leapfrogs=out.n_leapfrogs,
                   n_kernel_launches=out.n_kernel_launches)
        res["d2h_bytes"] = sum(v.nbytes for v in res.values() if isinstance(v, np.ndarray))
        return res

    def close(self):
        if self.h:
            self.lib.smc_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def philox_normals(seed, chain, step, n):
    out = np.empty(n)
    check(load().smc_philox_normals(seed, chain, step, 0, n, dptr(out)))
    return out


def philox_uniforms(seed, chain, step, n):
    out = np.empty(n)
    check(load().smc_philox_uniforms(seed, chain, step, 1, n, dptr(out)))
    return out


def plan_describe(program_bytes):
    """test hook: the planned micro-ops of one packed register program, as text lines (host only, no device needed)"""
    lib = load()
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib.smc_vm_plan_describe(program_bytes, len(program_bytes), buf, len(buf))
    check(n if n < 0 else 0)
    return buf.value.decode().splitlines()
