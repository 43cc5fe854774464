"""numpy interpreter of the sweep ISA (enoki_b200/csrc/ek_isa.h) for the CPU test-suite: executes the programs that the
planner + assembler produce (`enoki_b200.debug_program()`), so that scheduling, slot allocation, superinstruction
fusion, reduction phases and operand encoding are checked on machines without a GPU.  TEST INFRASTRUCTURE: it mirrors
what ek_sweep.cu does per element (32-bit value types only; 64-bit planes, gathers / scatters and the shared-memory
helpers raise Unsupported) and uses the C oracle for the operations numpy cannot round identically (fma, Cephes)."""
import ctypes

import numpy as np

F_ST, F_R64, F_HAS_B, F_HAS_C, F_HAS_A, F_NEG_A, F_ABS_A, F_STG, F_RACC = 1, 2, 4, 8, 0x80, 0x100, 0x200, 0x400, 0x4000
OP_NONE = 0xFFFF
SZ = ctypes.c_size_t
T_FLOAT32, T_INT32, T_UINT32, T_BOOL = 10, 5, 6, 12


class Unsupported(Exception):
    pass


def _P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Emulator:
    def __init__(self, oracle, arrays):
        """arrays: variable index -> numpy array (uint32 / int32 / float32 / bool data of evaluated inputs)."""
        self.oracle = oracle
        self.vars = {k: np.ascontiguousarray(v) for k, v in arrays.items()}

    # ---- helpers
    @staticmethod
    def _f(x): return x.view(np.float32)
    @staticmethod
    def _u(x): return np.ascontiguousarray(x).view(np.uint32)

    def _unary(self, which, x):
        xf = np.ascontiguousarray(self._f(x)); out = np.zeros_like(xf)
        self.oracle.or_unary_f32(which, _P(xf), _P(out), SZ(len(xf)))
        return self._u(out)

    def _fma(self, a, b, c):
        a, b, c = (np.ascontiguousarray(self._f(v)) for v in (a, b, c)); out = np.zeros_like(a)
        self.oracle.or_fma_f32(_P(a), _P(b), _P(c), _P(out), SZ(len(a)))
        return self._u(out)

    def _minmax(self, is_max, a, b):
        a, b = (np.ascontiguousarray(self._f(v)) for v in (a, b)); out = np.zeros_like(a)
        self.oracle.or_minmax_f32(int(is_max), _P(a), _P(b), _P(out), SZ(len(a)))
        return self._u(out)

    def run(self, program):
        names = program["ops"]
        for sw in program["sweeps"]:
            self._sweep(names, sw)

    def _sweep(self, names, sw):
        n = sw["n"]
        lits, argw = sw["lits"], list(sw["argw"])
        n_lit, n_arg = len(lits), len(argw)
        ptr = {}                                    # argw index -> variable
        for aw, var, _out in sw["ptr_fix"]:
            ptr[aw] = var
        # uniform pool = literals | argument words | scalar (lo, hi) pairs
        uni = [np.uint32(v) for v in lits] + [None] * n_arg
        for var, typ in sw["scalars"]:
            v = self.vars[var]
            if v.dtype.itemsize == 8:
                raise Unsupported("64-bit scalar")
            lo = np.uint32(1 if (typ == T_BOOL and v.reshape(-1)[0]) else v.reshape(-1).view(np.uint32 if v.dtype.itemsize == 4 else v.dtype)[0])
            uni += [lo, np.uint32(0)]
        staged = {}
        for var, unit, es in sw["staged"]:
            if es != 4:
                raise Unsupported("staged input of element size %d" % es)
            a = self.vars[var]
            assert a.size == n, (var, a.size, n)
            staged[unit] = self._u(a)
        out_type = {var: typ for var, _aw, _bytes, typ in sw["outputs"]}
        slots = {}
        state = {"R": np.zeros(n, np.uint32)}
        idx = np.arange(n, dtype=np.uint32)

        def bcast(v):
            return np.full(n, v, np.uint32)

        def fetch(code):
            if code == OP_NONE:
                raise AssertionError("fetch of an absent operand")
            if (code & 0xC000) == 0xC000:
                raise Unsupported("global operand")
            if code & 0x8000:
                v = uni[code & 0x3fff]
                if v is None:
                    raise AssertionError("argument word used as a value")
                return bcast(v)
            if code & 0x4000:
                return staged[code & 0x3fff].copy()
            if code not in slots:
                raise AssertionError("read of slot %d before it was written" % code)
            return slots[code].copy()

        def var_of_uniform(ui):
            aw = ui - n_lit
            assert 0 <= aw < n_arg and aw in ptr, ("uniform index is not a pointer argument", ui)
            return ptr[aw]

        def store_var(var, R):
            typ = out_type.get(var, T_UINT32)
            if typ == T_BOOL:
                self.vars[var] = (R & 1).astype(np.bool_)
            elif typ == T_FLOAT32:
                self.vars[var] = R.view(np.float32).copy()
            elif typ == T_INT32:
                self.vars[var] = R.view(np.int32).copy()
            elif typ == T_UINT32:
                self.vars[var] = R.copy()
            else:
                raise Unsupported("output type %d" % typ)

        def racc(slot, kc):
            kind, cls = kc & 0xff, (kc >> 8) & 0xff
            acc = slots[slot]
            R = state["R"]
            with np.errstate(all="ignore"):
                if cls == 0:
                    a, r = self._f(acc), self._f(R)
                    res = {0: a + r, 1: a * r}.get(kind)
                    if res is None:
                        res = self._f(self._minmax(kind == 3, R, acc))       # red_combine: min_x86(y, x)
                elif cls in (1, 2):
                    a = acc.view(np.int32 if cls == 1 else np.uint32); r = R.view(a.dtype)
                    res = {0: a + r, 1: a * r, 2: np.minimum(a, r), 3: np.maximum(a, r)}[kind]
                else:
                    raise Unsupported("64-bit reduction")
            slots[slot] = self._u(np.ascontiguousarray(res))

        def rfin(B, imm, dst):
            kind, cls = imm & 0xff, (imm >> 8) & 0xff
            if cls > 2:
                raise Unsupported("64-bit reduction")
            if cls == 0:
                x = self._f(B).astype(np.float64)
                r = {0: x.sum(), 1: np.prod(x), 2: x.min(), 3: x.max()}[kind]
                val = np.array([r], np.float32)
            else:
                dt = np.int32 if cls == 1 else np.uint32
                x = B.view(dt)
                if kind == 0: r = dt(int(x.astype(np.int64).sum()) & 0xffffffff) if dt is np.uint32 else np.int64(x.astype(np.int64).sum()).astype(np.int32)
                elif kind == 1:
                    acc = 1
                    for v in x.tolist(): acc = (acc * v) & 0xffffffff
                    r = np.uint32(acc).astype(dt) if dt is np.uint32 else np.uint32(acc).view(np.int32)
                else: r = x.min() if kind == 2 else x.max()
                val = np.array([r]).astype(dt)
            self.vars[var_of_uniform(dst)] = val

        def execute(ins):
            op, flags, dst, cb, cc, ca, imm = ins
            name = names[op]
            if flags & F_R64:
                raise Unsupported("64-bit value")
            if flags & F_HAS_A:
                state["R"] = fetch(ca)
            if flags & F_ABS_A:
                state["R"] = state["R"] & np.uint32(0x7fffffff)
            if flags & F_NEG_A:
                state["R"] = state["R"] ^ np.uint32(0x80000000)
            R = state["R"]
            B = fetch(cb) if flags & F_HAS_B else None
            C = fetch(cc) if flags & F_HAS_C else None
            f = self._f
            with np.errstate(all="ignore"):
                if name == "NOP": pass
                elif name == "ADD_F32": R = self._u(f(R) + f(B))
                elif name == "SUB_F32": R = self._u(f(R) - f(B))
                elif name == "SUBR_F32": R = self._u(f(B) - f(R))
                elif name == "MUL_F32": R = self._u(f(R) * f(B))
                elif name == "DIV_F32": R = self._u(f(R) / f(B))
                elif name == "DIVR_F32": R = self._u(f(B) / f(R))
                elif name == "FMA_F32": R = self._fma(R, B, C)
                elif name == "FMAC_F32": R = self._fma(B, C, R)
                elif name in ("MIN_F32", "MAX_F32"): R = self._minmax(name == "MAX_F32", R, B)
                elif name in ("MINR_F32", "MAXR_F32"): R = self._minmax(name == "MAXR_F32", B, R)
                elif name == "ABS_F32": R = R & np.uint32(0x7fffffff)
                elif name == "NEG_F32": R = R ^ np.uint32(0x80000000)
                elif name == "SQRT_F32": R = self._u(np.sqrt(f(R)))
                elif name == "EXP_F32": R = self._unary(2, R)
                elif name == "LOG_F32": R = self._unary(3, R)
                elif name == "SIN_F32": R = self._unary(0, R)
                elif name == "COS_F32": R = self._unary(1, R)
                elif name == "FLOOR_F32": R = self._u(np.floor(f(R)))
                elif name == "CEIL_F32": R = self._u(np.ceil(f(R)))
                elif name == "ROUND_F32": R = self._u(np.rint(f(R)))
                elif name == "TRUNC_F32": R = self._u(np.trunc(f(R)))
                elif name in ("LT_F32", "LE_F32", "GT_F32", "GE_F32", "EQ_F32", "NE_F32"):
                    fn = {"LT": np.less, "LE": np.less_equal, "GT": np.greater, "GE": np.greater_equal, "EQ": np.equal, "NE": np.not_equal}[name[:2]]
                    R = fn(f(R), f(B)).astype(np.uint32)
                elif name == "ADD_I32": R = R + B
                elif name == "SUB_I32": R = R - B
                elif name == "SUBR_I32": R = B - R
                elif name == "MUL_I32": R = R * B
                elif name == "MAD_I32": R = R * B + C
                elif name == "MADC_I32": R = B * C + R
                elif name == "NEG_I32": R = np.uint32(0) - R
                elif name == "NOT_32": R = ~R
                elif name == "AND_32": R = R & B
                elif name == "OR_32": R = R | B
                elif name == "XOR_32": R = R ^ B
                elif name == "SHL_32": R = np.where(B >= 32, np.uint32(0), R << (B & np.uint32(31)))
                elif name == "SHR_U32": R = np.where(B >= 32, np.uint32(0), R >> (B & np.uint32(31)))
                elif name == "SHR_I32": R = (R.view(np.int32) >> np.minimum(B, np.uint32(31)).astype(np.int32)).view(np.uint32)
                elif name in ("MIN_U32", "MAX_U32"): R = (np.minimum if name == "MIN_U32" else np.maximum)(R, B)
                elif name in ("MIN_I32", "MAX_I32"): R = (np.minimum if name == "MIN_I32" else np.maximum)(R.view(np.int32), B.view(np.int32)).view(np.uint32)
                elif name in ("LT_U32", "LE_U32", "GT_U32", "GE_U32"):
                    fn = {"LT": np.less, "LE": np.less_equal, "GT": np.greater, "GE": np.greater_equal}[name[:2]]
                    R = fn(R, B).astype(np.uint32)
                elif name in ("LT_I32", "LE_I32", "GT_I32", "GE_I32"):
                    fn = {"LT": np.less, "LE": np.less_equal, "GT": np.greater, "GE": np.greater_equal}[name[:2]]
                    R = fn(R.view(np.int32), B.view(np.int32)).astype(np.uint32)
                elif name == "EQ_32": R = (R == B).astype(np.uint32)
                elif name == "NE_32": R = (R != B).astype(np.uint32)
                elif name == "NOT_B": R = R ^ np.uint32(1)
                elif name == "NEZ_32": R = (R != 0).astype(np.uint32)
                elif name == "SEL_M_32": R = np.where(R != 0, B, C)
                elif name == "SEL_T_32": R = np.where(B != 0, R, C)
                elif name == "SEL_F_32": R = np.where(B != 0, C, R)
                elif name == "LOAD_32": R = B
                elif name == "INDEX": R = idx.copy()
                elif name == "CVT_F32_U32":
                    x = f(R)
                    if imm != 0:
                        raise Unsupported("rounded float->uint")
                    ok = (x > -9.2233720368547758e18) & (x < 9.2233720368547758e18)
                    R = np.where(ok, np.trunc(np.where(ok, x, 0)).astype(np.int64).astype(np.uint64) & np.uint64(0xffffffff), 0).astype(np.uint32)
                elif name == "CVT_F32_I32":
                    x = f(R)
                    r = {0: np.trunc, 1: np.floor, 2: np.ceil, 3: np.rint}[imm](x)
                    ok = (x >= -2147483648.0) & (x < 2147483648.0)
                    R = np.where(ok, np.where(ok, r, 0).astype(np.int64), np.int64(-2147483648)).astype(np.int32).view(np.uint32)
                elif name == "CVT_I32_F32": R = self._u(R.view(np.int32).astype(np.float32))
                elif name == "CVT_U32_F32": R = self._u(R.astype(np.float32))
                elif name == "LDG_32": R = self._u(self.vars[var_of_uniform(imm)]).copy()
                elif name == "ST_32": store_var(var_of_uniform(imm), R)
                elif name == "RACC": state["R"] = R; racc(dst, imm); return
                elif name == "RFIN": rfin(B, imm, dst); return
                else:
                    raise Unsupported(name)
            state["R"] = np.ascontiguousarray(R).astype(np.uint32, copy=False)
            if flags & F_RACC:
                racc(dst, ca)
            if flags & F_STG:
                store_var(var_of_uniform(imm), state["R"])
            if flags & F_ST:
                slots[dst] = state["R"].copy()

        for ins in sw["init"]:
            execute(ins)
        for ins in sw["body"]:
            execute(ins)
        for ins in sw["fini"]:
            execute(ins)
