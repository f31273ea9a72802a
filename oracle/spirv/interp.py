"""Minimal SPIR-V interpreter for the reference's SHIPPED DDGI compute shaders — TEST INFRASTRUCTURE.

Purpose: pin the CPU oracle against the reference's own compiled binaries
(`/root/reference/Assets/shaders/spv/DDGI/{GISDFRays,IrradianceProbeUpdate,DepthProbeUpdate,IrradianceBorderUpdate,
DepthBorderUpdate}.comp.spv`).  The interpreter executes the SPIR-V invocation by invocation (workgroups in lock-step at
OpControlBarrier) on small inputs; `tests/golden/make_spirv_golden.py` stores the results as fixtures.

It also executes the shipped binaries of the rows either side of the path (`DDGI/SampleProbe.comp`, `SDF/{SDFRasterizeModel,
SDFRasterizeModelNoRead,GlobalSDFMipmap,SDFCulling,SDFReflection,SDFShadow}.comp` and the fragment shaders `SDF/{SDFDeferredLight,
SDFAtlasIndirectLight}.frag`, one invocation per fragment with the stage inputs set by the caller): tests/golden/make_spirv_golden_*.py.

Scope: exactly the ~80 core opcodes and 18 GLSL.std.450 instructions those fourteen modules use (opcode numbers from the
public SPIR-V 1.5 specification).  Anything else raises NotImplementedError.

Numerics: binary32 via numpy.float32 scalars, one rounding per SPIR-V instruction, NO contraction.  Operations whose
precision Vulkan leaves to the implementation follow the repo's numerics contract: sin/cos/pow through
binary64, normalize = v * (1/sqrt(dot)), MatrixInverse = cofactor expansion, Cross / Reflect with every product and sum rounded, FMin/FMax/FClamp select-based, texture
filtering with fp32 weights and nested lerps, image stores RTNE to fp16.
"""
from __future__ import annotations

import math
import struct

import numpy as np

F = np.float32
M32 = 0xFFFFFFFF

OP = {
    1: "Undef", 3: "Source", 4: "SourceExtension", 5: "Name", 6: "MemberName", 10: "Extension", 11: "ExtInstImport", 12: "ExtInst",
    14: "MemoryModel", 15: "EntryPoint", 16: "ExecutionMode", 17: "Capability", 19: "TypeVoid", 20: "TypeBool", 21: "TypeInt",
    22: "TypeFloat", 23: "TypeVector", 24: "TypeMatrix", 25: "TypeImage", 26: "TypeSampler", 27: "TypeSampledImage", 28: "TypeArray",
    29: "TypeRuntimeArray", 30: "TypeStruct", 32: "TypePointer", 33: "TypeFunction", 41: "ConstantTrue", 42: "ConstantFalse",
    43: "Constant", 44: "ConstantComposite", 46: "ConstantNull", 54: "Function", 55: "FunctionParameter", 56: "FunctionEnd",
    57: "FunctionCall", 59: "Variable", 61: "Load", 62: "Store", 65: "AccessChain", 66: "InBoundsAccessChain", 71: "Decorate",
    72: "MemberDecorate", 77: "VectorExtractDynamic", 79: "VectorShuffle", 80: "CompositeConstruct", 81: "CompositeExtract", 82: "CompositeInsert",
    83: "CopyObject", 87: "ImageSampleImplicitLod", 88: "ImageSampleExplicitLod", 95: "ImageFetch", 96: "ImageGather", 98: "ImageRead", 99: "ImageWrite",
    100: "Image", 103: "ImageQuerySizeLod", 110: "ConvertFToS", 109: "ConvertFToU", 111: "ConvertSToF", 112: "ConvertUToF", 124: "Bitcast", 126: "SNegate",
    127: "FNegate", 128: "IAdd", 129: "FAdd", 130: "ISub", 131: "FSub", 132: "IMul", 133: "FMul", 134: "UDiv", 135: "SDiv",
    136: "FDiv", 137: "UMod", 138: "SRem", 139: "SMod", 141: "FMod", 142: "VectorTimesScalar", 143: "MatrixTimesScalar",
    144: "VectorTimesMatrix", 145: "MatrixTimesVector", 146: "MatrixTimesMatrix", 148: "Dot", 154: "Any", 155: "All",
    164: "LogicalEqual", 165: "LogicalNotEqual", 166: "LogicalOr", 167: "LogicalAnd", 168: "LogicalNot", 169: "Select",
    170: "IEqual", 171: "INotEqual", 172: "UGreaterThan", 173: "SGreaterThan", 174: "UGreaterThanEqual", 175: "SGreaterThanEqual",
    176: "ULessThan", 177: "SLessThan", 178: "ULessThanEqual", 179: "SLessThanEqual", 180: "FOrdEqual", 182: "FOrdNotEqual", 183: "FUnordNotEqual",
    184: "FOrdLessThan", 186: "FOrdGreaterThan", 188: "FOrdLessThanEqual", 190: "FOrdGreaterThanEqual", 194: "ShiftRightLogical",
    195: "ShiftRightArithmetic", 196: "ShiftLeftLogical", 197: "BitwiseOr", 198: "BitwiseXor", 199: "BitwiseAnd", 200: "Not",
    224: "ControlBarrier", 225: "MemoryBarrier", 234: "AtomicIAdd", 241: "AtomicOr", 245: "Phi", 246: "LoopMerge", 247: "SelectionMerge", 248: "Label", 249: "Branch",
    250: "BranchConditional", 252: "Kill", 253: "Return", 254: "ReturnValue", 255: "Unreachable", 400: "CopyLogical",
}
GLSL = {4: "FAbs", 8: "Floor", 10: "Fract", 13: "Sin", 14: "Cos", 26: "Pow", 31: "Sqrt", 32: "InverseSqrt", 34: "MatrixInverse",
        37: "FMin", 38: "UMin", 39: "SMin", 40: "FMax", 41: "UMax", 42: "SMax", 43: "FClamp", 44: "UClamp", 45: "SClamp", 46: "FMix", 48: "Step",
        66: "Length", 67: "Distance", 68: "Cross", 69: "Normalize", 71: "Reflect"}
DEC_BUILTIN, DEC_BINDING, DEC_SET = 11, 33, 34
BUILTIN = {24: "NumWorkgroups", 25: "WorkgroupSize", 26: "WorkgroupId", 27: "LocalInvocationId", 28: "GlobalInvocationId", 29: "LocalInvocationIndex"}
SC_FUNCTION, SC_WORKGROUP = 7, 4


def s32(x):
    x &= M32
    return x - (1 << 32) if x & 0x80000000 else x


def sdiv(p, q):
    """OpSDiv: signed division truncating toward zero."""
    p, q = s32(p), s32(q)
    r = abs(p) // abs(q)
    return (r if (p < 0) == (q < 0) else -r) & M32


def f2h(f):
    """binary32 -> binary16 bits, round to nearest even (numpy implements IEEE RTNE)."""
    with np.errstate(over="ignore"):
        return int(np.float32(f).astype(np.float16).view(np.uint16))


def h2f(h):
    return F(np.uint16(h).view(np.float16))


# ----------------------------------------------------------------------------------------------------------------------
# Resources
# ----------------------------------------------------------------------------------------------------------------------
def gmin(x, y):
    return y if y < x else x


def gmax(x, y):
    return y if x < y else x


def gclamp(x, lo, hi):
    return gmin(gmax(x, lo), hi)


def lerp(a, b, t):
    return F(a + F(t * F(b - a)))


class Texture3D:
    """R16F sampler3D, linear filter, clamp-to-edge."""

    def __init__(self, data_f16):  # [z][y][x]
        self.d = np.ascontiguousarray(data_f16, dtype=np.float16)
        self.D, self.H, self.W = self.d.shape
        self.taps = 0

    def sample(self, c):
        self.taps += 1
        x = F(F(c[0] * F(self.W)) - F(0.5)); y = F(F(c[1] * F(self.H)) - F(0.5)); z = F(F(c[2] * F(self.D)) - F(0.5))
        fx, fy, fz = F(math.floor(x)), F(math.floor(y)), F(math.floor(z))
        ax, ay, az = F(x - fx), F(y - fy), F(z - fz)
        ix, iy, iz = int(fx), int(fy), int(fz)
        cl = lambda v, n: min(max(v, 0), n - 1)
        x0, x1, y0, y1, z0, z1 = cl(ix, self.W), cl(ix + 1, self.W), cl(iy, self.H), cl(iy + 1, self.H), cl(iz, self.D), cl(iz + 1, self.D)
        t = lambda xx, yy, zz: F(self.d[zz, yy, xx])
        c00 = lerp(t(x0, y0, z0), t(x1, y0, z0), ax); c10 = lerp(t(x0, y1, z0), t(x1, y1, z0), ax)
        c01 = lerp(t(x0, y0, z1), t(x1, y0, z1), ax); c11 = lerp(t(x0, y1, z1), t(x1, y1, z1), ax)
        r = lerp(lerp(c00, c10, ay), lerp(c01, c11, ay), az)
        return [r, F(0), F(0), F(1)]


class Texture2D:
    """sampler2D used through textureGather / texelFetch.  data: float array [h][w][channels] (fp16 or fp32)."""

    def __init__(self, data, repeat):
        self.d = np.ascontiguousarray(data)
        if self.d.ndim == 2:
            self.d = self.d[:, :, None]
        self.H, self.W, self.C = self.d.shape
        self.repeat = repeat

    def _wrap(self, i, n):
        return i % n if self.repeat else min(max(i, 0), n - 1)

    def gather(self, c, comp):
        i0 = int(math.floor(F(F(c[0] * F(self.W)) - F(0.5)))); j0 = int(math.floor(F(F(c[1] * F(self.H)) - F(0.5))))
        i1, j1 = i0 + 1, j0 + 1
        i0, i1, j0, j1 = self._wrap(i0, self.W), self._wrap(i1, self.W), self._wrap(j0, self.H), self._wrap(j1, self.H)
        g = lambda i, j: F(self.d[j, i, comp]) if comp < self.C else F(0)
        return [g(i0, j1), g(i1, j1), g(i1, j0), g(i0, j0)]

    def fetch(self, c):
        y, x = s32(c[1]), s32(c[0])
        if not (0 <= y < self.H and 0 <= x < self.W):  # out-of-bounds texelFetch: robust-buffer-access style zeros
            return [F(0), F(0), F(0), F(0)]
        px = self.d[y, x]
        out = [F(px[k]) if k < self.C else F(0) for k in range(3)]
        out.append(F(px[3]) if self.C > 3 else F(1))
        return out

    def sample(self, c):
        """textureLod(sampler2D, uv, 0): bilinear with fp32 weights, nested lerp x then y."""
        x = F(F(c[0] * F(self.W)) - F(0.5)); y = F(F(c[1] * F(self.H)) - F(0.5))
        fx, fy = F(math.floor(x)), F(math.floor(y))
        ax, ay = F(x - fx), F(y - fy)
        x0, x1, y0, y1 = self._wrap(int(fx), self.W), self._wrap(int(fx) + 1, self.W), self._wrap(int(fy), self.H), self._wrap(int(fy) + 1, self.H)
        out = []
        for k in range(4):
            if k < self.C:
                t = lambda xx, yy: F(self.d[yy, xx, k])
                out.append(lerp(lerp(t(x0, y0), t(x1, y0), ax), lerp(t(x0, y1), t(x1, y1), ax), ay))
            else:
                out.append(F(0) if k < 3 else F(1))
        return out


class TextureCube:
    """samplerCube, bilinear inside the face, clamp at face edges; faces [+x,-x,+y,-y,+z,-z][n][n][4]."""

    def __init__(self, faces):
        self.f = None if faces is None else np.ascontiguousarray(faces, dtype=np.float16)

    def sample(self, d):
        if self.f is None:
            return [F(0), F(0), F(0), F(1)]
        ax, ay, az = abs(d[0]), abs(d[1]), abs(d[2])
        if az >= ax and az >= ay:
            face, sc, tc, ma = (4 if d[2] >= 0 else 5), (d[0] if d[2] >= 0 else F(-d[0])), F(-d[1]), az
        elif ay >= ax:
            face, sc, tc, ma = (2 if d[1] >= 0 else 3), d[0], (d[2] if d[1] >= 0 else F(-d[2])), ay
        else:
            face, sc, tc, ma = (0 if d[0] >= 0 else 1), (F(-d[2]) if d[0] >= 0 else d[2]), F(-d[1]), ax
        u = F(F(F(0.5) * F(sc / ma)) + F(0.5)); v = F(F(F(0.5) * F(tc / ma)) + F(0.5))
        N = self.f.shape[1]
        x = F(F(u * F(N)) - F(0.5)); y = F(F(v * F(N)) - F(0.5))
        fx, fy = F(math.floor(x)), F(math.floor(y))
        axw, ayw = F(x - fx), F(y - fy)
        cl = lambda q: min(max(q, 0), N - 1)
        x0, x1, y0, y1 = cl(int(fx)), cl(int(fx) + 1), cl(int(fy)), cl(int(fy) + 1)
        out = []
        for ch in range(4):
            t = lambda xx, yy: F(self.f[face, yy, xx, ch])
            out.append(lerp(lerp(t(x0, y0), t(x1, y0), axw), lerp(t(x0, y1), t(x1, y1), axw), ayw))
        return out


class Texture3DMips:
    """sampler3D with a mip chain (mesh distance fields): textureLod at an INTEGER lod = trilinear inside that level,
    REPEAT addressing (the reference's Texture3D default wrap).  levels: list of fp16 arrays [d][h][w]."""

    def __init__(self, levels):
        self.levels = [np.ascontiguousarray(l, dtype=np.float16) for l in levels]
        self.taps = 0

    def sample(self, c, lod=0.0):
        assert float(lod) == int(lod), "fractional LOD is not modelled"
        d = self.levels[int(lod)]
        D, H, W = d.shape
        self.taps += 1
        x = F(F(c[0] * F(W)) - F(0.5)); y = F(F(c[1] * F(H)) - F(0.5)); z = F(F(c[2] * F(D)) - F(0.5))
        fx, fy, fz = F(math.floor(x)), F(math.floor(y)), F(math.floor(z))
        ax, ay, az = F(x - fx), F(y - fy), F(z - fz)
        ix, iy, iz = int(fx), int(fy), int(fz)
        x0, x1, y0, y1, z0, z1 = ix % W, (ix + 1) % W, iy % H, (iy + 1) % H, iz % D, (iz + 1) % D
        t = lambda xx, yy, zz: F(d[zz, yy, xx])
        c00 = lerp(t(x0, y0, z0), t(x1, y0, z0), ax); c10 = lerp(t(x0, y1, z0), t(x1, y1, z0), ax)
        c01 = lerp(t(x0, y0, z1), t(x1, y0, z1), ax); c11 = lerp(t(x0, y1, z1), t(x1, y1, z1), ax)
        r = lerp(lerp(c00, c10, ay), lerp(c01, c11, ay), az)
        return [r, F(0), F(0), F(1)]


class StorageImage3D:
    """image3D r16f: uint16 array [d][h][w]; out-of-bounds accesses are discarded / read zero (robust image access)."""

    def __init__(self, bits_u16):
        self.b = bits_u16

    def _in(self, c):
        z, y, x = s32(c[2]), s32(c[1]), s32(c[0])
        D, H, W = self.b.shape
        return (z, y, x) if (0 <= z < D and 0 <= y < H and 0 <= x < W) else None

    def write(self, c, texel):
        i = self._in(c)
        if i is not None:
            self.b[i] = f2h(texel[0])

    def read(self, c):
        i = self._in(c)
        return [h2f(self.b[i]) if i is not None else F(0), F(0), F(0), F(1)]


class StorageImage:
    """image2D with an fp16 format (rgba16f / rg16f): uint16 array [h][w][channels]."""

    def __init__(self, bits_u16):
        self.b = bits_u16
        self.C = bits_u16.shape[2]

    def write(self, c, texel):
        for k in range(self.C):
            self.b[s32(c[1]), s32(c[0]), k] = f2h(texel[k])

    def read(self, c):
        px = self.b[s32(c[1]), s32(c[0])]
        out = [h2f(px[k]) if k < self.C else F(0) for k in range(3)]
        out.append(h2f(px[3]) if self.C > 3 else F(1))
        return out

    # a storage image bound as sampler2D for texelFetch (the blend reads the ray buffers / previous atlases that way)
    def fetch(self, c):
        return self.read(c)


# ----------------------------------------------------------------------------------------------------------------------
# Module
# ----------------------------------------------------------------------------------------------------------------------
class Ptr:
    __slots__ = ("cell", "path")

    def __init__(self, cell, path=()):
        self.cell, self.path = cell, path

    def load(self):
        v = self.cell[0]
        for i in self.path:
            v = v[i]
        return copyv(v)

    def store(self, val):
        if not self.path:
            self.cell[0] = copyv(val)
            return
        v = self.cell[0]
        for i in self.path[:-1]:
            v = v[i]
        v[self.path[-1]] = copyv(val)


def copyv(v):
    return [copyv(x) for x in v] if type(v) is list else v


class Barrier:
    pass


class Module:
    def __init__(self, path):
        raw = open(path, "rb").read()
        w = struct.unpack("<%dI" % (len(raw) // 4), raw)
        assert w[0] == 0x07230203, "not SPIR-V"
        self.types, self.consts, self.names, self.decor, self.globals_, self.functions = {}, {}, {}, {}, {}, {}
        self.entry, self.local_size, self.ext = None, (1, 1, 1), {}
        i, cur = 5, None
        while i < len(w):
            wc, op = w[i] >> 16, w[i] & 0xFFFF
            a = w[i + 1:i + wc]
            name = OP.get(op)
            if name is None:
                raise NotImplementedError(f"SPIR-V opcode {op}")
            i += wc
            if name == "Name":
                self.names[a[0]] = self._str(a[1:])
            elif name == "Decorate":
                self.decor.setdefault(a[0], {})[a[1]] = a[2] if len(a) > 2 else True
            elif name == "ExtInstImport":
                self.ext[a[0]] = self._str(a[1:])
            elif name == "EntryPoint":
                self.entry = a[1]
            elif name == "ExecutionMode":
                if a[1] == 17:  # LocalSize
                    self.local_size = (a[2], a[3], a[4])
            elif name.startswith("Type"):
                self.types[a[0]] = (name[4:],) + tuple(a[1:])
            elif name in ("Constant", "ConstantTrue", "ConstantFalse", "ConstantComposite", "ConstantNull"):
                self.consts[a[1]] = self._const(name, a)
            elif name == "Variable" and cur is None:
                self.globals_[a[1]] = (a[0], a[2], a[3] if len(a) > 3 else None)
            elif name == "Function":
                cur = {"id": a[1], "params": [], "blocks": {}, "order": [], "rtype": a[0]}
                self.functions[a[1]] = cur
                blk = None
            elif name == "FunctionParameter":
                cur["params"].append(a[1])
            elif name == "Label":
                blk = []
                cur["blocks"][a[0]] = blk
                cur["order"].append(a[0])
            elif name == "FunctionEnd":
                cur = None
            elif cur is not None:
                blk.append((name, a))
        self.storage = {}  # global id -> cell

    @staticmethod
    def _str(words):
        b = b"".join(struct.pack("<I", x) for x in words)
        return b.split(b"\0")[0].decode()

    def _const(self, name, a):
        t = self.types[a[0]]
        if name == "ConstantTrue":
            return True
        if name == "ConstantFalse":
            return False
        if name == "ConstantComposite":
            return [copyv(self.consts[c]) for c in a[2:]]
        if name == "ConstantNull":
            return self.zero(a[0])
        if t[0] == "Float":
            return F(struct.unpack("<f", struct.pack("<I", a[2]))[0])
        return a[2] & M32

    def zero(self, tid):
        t = self.types[tid]
        k = t[0]
        if k == "Float":
            return F(0)
        if k == "Int":
            return 0
        if k == "Bool":
            return False
        if k in ("Vector", "Matrix"):
            return [self.zero(t[1]) for _ in range(t[2])]
        if k == "Array":
            return [self.zero(t[1]) for _ in range(self.consts[t[2]])]
        if k == "Struct":
            return [self.zero(m) for m in t[1:]]
        if k == "RuntimeArray":
            return []
        return None

    def binding(self, gid):
        d = self.decor.get(gid, {})
        return d.get(DEC_SET), d.get(DEC_BINDING)

    def global_by_binding(self, set_, binding):
        for gid in self.globals_:
            if self.binding(gid) == (set_, binding):
                return gid
        return None

    def global_by_storage(self, sc):
        return [g for g, (_, s, _) in self.globals_.items() if s == sc]


class Invocation:
    """One shader invocation as a generator: yields Barrier at OpControlBarrier."""

    def __init__(self, mod: Module, shared: dict, builtins: dict):
        self.m, self.shared, self.builtins = mod, shared, builtins
        self.g = {}
        for gid, (tid, sc, init) in mod.globals_.items():
            if gid in mod.storage:
                self.g[gid] = Ptr(mod.storage[gid])
            elif sc == SC_WORKGROUP:
                self.g[gid] = Ptr(shared.setdefault(gid, [mod.zero(mod.types[tid][2])]))
            elif DEC_BUILTIN in mod.decor.get(gid, {}):
                self.g[gid] = Ptr([copyv(builtins[BUILTIN[mod.decor[gid][DEC_BUILTIN]]])])
            else:  # Private / Output ...
                self.g[gid] = Ptr([copyv(mod.consts[init]) if init is not None else mod.zero(mod.types[tid][2])])
        self.count = 0

    def run(self):
        yield from self.call(self.m.functions[self.m.entry], [])

    def val(self, env, i):
        if i in env:
            return env[i]
        c = self.m.consts.get(i)
        if c is not None or i in self.m.consts:
            return c
        return self.g[i]

    def call(self, fn, args):
        m = self.m
        env = dict(zip(fn["params"], args))
        V = lambda i: self.val(env, i)
        label, prev = fn["order"][0], None
        while True:
            nxt = None
            for name, a in fn["blocks"][label]:
                self.count += 1
                if name == "Load":
                    p = V(a[2])
                    env[a[1]] = p.load() if isinstance(p, Ptr) else p
                elif name == "Store":
                    V(a[0]).store(V(a[1]))
                elif name in ("AccessChain", "InBoundsAccessChain"):
                    p = V(a[2])
                    env[a[1]] = Ptr(p.cell, p.path + tuple(s32(V(x)) for x in a[3:]))
                elif name == "Variable":
                    tid, init = a[0], (a[3] if len(a) > 3 else None)
                    env[a[1]] = Ptr([copyv(m.consts[init]) if init is not None else m.zero(m.types[tid][2])])
                elif name == "FunctionCall":
                    r = yield from self.call(m.functions[a[2]], [V(x) for x in a[3:]])
                    env[a[1]] = r
                elif name == "ReturnValue":
                    return copyv(V(a[0]))
                elif name == "Return":
                    return None
                elif name == "Branch":
                    nxt = a[0]
                elif name == "BranchConditional":
                    nxt = a[1] if V(a[0]) else a[2]
                elif name in ("SelectionMerge", "LoopMerge"):
                    pass
                elif name == "Phi":
                    for k in range(2, len(a), 2):
                        if a[k + 1] == prev:
                            env[a[1]] = copyv(V(a[k]))
                elif name == "ControlBarrier":
                    yield Barrier
                elif name == "MemoryBarrier":
                    pass
                elif name == "ImageWrite":
                    V(a[0]).write(V(a[1]), V(a[2]))
                elif name == "AtomicIAdd":  # result type, result id, pointer, scope, semantics, value; invocations run one at a time
                    ptr = V(a[2])
                    old = ptr.load()
                    ptr.store((old + V(a[5])) & M32)
                    env[a[1]] = old
                elif name == "AtomicOr":
                    ptr = V(a[2])
                    old = ptr.load()
                    ptr.store((old | V(a[5])) & M32)
                    env[a[1]] = old
                else:
                    env[a[1]] = self.alu(name, a, V)
            prev, label = label, nxt
            if nxt is None:
                raise RuntimeError("block fell through")

    # ------------------------------------------------------------------------------------------------------------------
    def alu(self, name, a, V):
        m = self.m
        rt = m.types[a[0]]
        x = lambda k: V(a[k])
        vec = lambda f, *ops: [f(*e) for e in zip(*ops)] if type(ops[0]) is list else f(*ops)

        if name == "FAdd": return vec(lambda p, q: F(p + q), x(2), x(3))
        if name == "FSub": return vec(lambda p, q: F(p - q), x(2), x(3))
        if name == "FMul": return vec(lambda p, q: F(p * q), x(2), x(3))
        if name == "FDiv":
            with np.errstate(divide="ignore", invalid="ignore"):
                return vec(lambda p, q: F(np.divide(p, q)), x(2), x(3))
        if name == "FNegate": return vec(lambda p: F(-p), x(2))
        if name == "IAdd": return vec(lambda p, q: (p + q) & M32, x(2), x(3))
        if name == "ISub": return vec(lambda p, q: (p - q) & M32, x(2), x(3))
        if name == "IMul": return vec(lambda p, q: (p * q) & M32, x(2), x(3))
        if name == "SDiv": return vec(sdiv, x(2), x(3))
        if name == "FMod": return vec(lambda p, q: F(p - F(q * F(math.floor(F(np.divide(p, q)))))), x(2), x(3))
        if name == "UDiv": return vec(lambda p, q: (p // q) & M32, x(2), x(3))
        if name == "SMod": return vec(lambda p, q: (s32(p) % s32(q)) & M32, x(2), x(3))  # sign follows the divisor, like Python
        if name == "UMod": return vec(lambda p, q: (p % q) & M32, x(2), x(3))
        if name == "SNegate": return vec(lambda p: (-p) & M32, x(2))
        if name == "VectorTimesScalar":
            s = x(3)
            return [F(e * s) for e in x(2)]
        if name == "MatrixTimesVector":
            mat, v = x(2), x(3)
            out = []
            for r in range(len(mat[0])):
                acc = F(mat[0][r] * v[0])
                for c in range(1, len(v)):
                    acc = F(acc + F(mat[c][r] * v[c]))
                out.append(acc)
            return out
        if name == "Dot":
            p, q = x(2), x(3)
            acc = F(p[0] * q[0])
            for k in range(1, len(p)):
                acc = F(acc + F(p[k] * q[k]))
            return acc
        if name == "ShiftRightLogical": return vec(lambda p, q: (p & M32) >> (q & 31), x(2), x(3))
        if name == "ShiftRightArithmetic": return vec(lambda p, q: (s32(p) >> (q & 31)) & M32, x(2), x(3))
        if name == "ShiftLeftLogical": return vec(lambda p, q: (p << (q & 31)) & M32, x(2), x(3))
        if name == "BitwiseXor": return vec(lambda p, q: (p ^ q) & M32, x(2), x(3))
        if name == "BitwiseOr": return vec(lambda p, q: (p | q) & M32, x(2), x(3))
        if name == "BitwiseAnd": return vec(lambda p, q: (p & q) & M32, x(2), x(3))
        if name == "Not": return vec(lambda p: (~p) & M32, x(2))
        cmp = {"ULessThan": lambda p, q: p < q, "ULessThanEqual": lambda p, q: p <= q, "UGreaterThan": lambda p, q: p > q,
               "UGreaterThanEqual": lambda p, q: p >= q, "SLessThan": lambda p, q: s32(p) < s32(q), "SLessThanEqual": lambda p, q: s32(p) <= s32(q),
               "SGreaterThan": lambda p, q: s32(p) > s32(q), "SGreaterThanEqual": lambda p, q: s32(p) >= s32(q),
               "IEqual": lambda p, q: (p & M32) == (q & M32), "INotEqual": lambda p, q: (p & M32) != (q & M32),
               "FUnordNotEqual": lambda p, q: bool(p != q) or bool(np.isnan(p) or np.isnan(q)),
               "FOrdEqual": lambda p, q: bool(p == q), "FOrdNotEqual": lambda p, q: bool(p != q) and not (np.isnan(p) or np.isnan(q)),
               "FOrdLessThan": lambda p, q: bool(p < q), "FOrdGreaterThan": lambda p, q: bool(p > q),
               "FOrdLessThanEqual": lambda p, q: bool(p <= q), "FOrdGreaterThanEqual": lambda p, q: bool(p >= q),
               "LogicalOr": lambda p, q: p or q, "LogicalAnd": lambda p, q: p and q, "LogicalEqual": lambda p, q: p == q,
               "LogicalNotEqual": lambda p, q: p != q}
        if name in cmp: return vec(cmp[name], x(2), x(3))
        if name == "LogicalNot": return vec(lambda p: not p, x(2))
        if name == "Any": return any(x(2))
        if name == "All": return all(x(2))
        if name == "Select":
            c, p, q = x(2), x(3), x(4)
            return [pp if cc else qq for cc, pp, qq in zip(c, p, q)] if type(c) is list else (copyv(p) if c else copyv(q))
        if name == "ConvertSToF": return vec(lambda p: F(s32(p)), x(2))
        if name == "ConvertUToF": return vec(lambda p: F(p & M32), x(2))
        if name == "ConvertFToS": return vec(lambda p: int(math.trunc(float(p))) & M32, x(2))
        if name == "ConvertFToU": return vec(lambda p: int(math.trunc(float(p))) & M32, x(2))
        if name == "Bitcast":
            def bc(p):
                if rt[0] == "Float" or (rt[0] == "Vector" and m.types[rt[1]][0] == "Float"):
                    return p if isinstance(p, np.floating) else F(struct.unpack("<f", struct.pack("<I", p & M32))[0])
                return struct.unpack("<I", struct.pack("<f", float(p)))[0] if isinstance(p, np.floating) else p & M32
            return vec(bc, x(2))
        if name == "VectorExtractDynamic": return copyv(x(2)[x(3) & M32])
        if name == "CompositeExtract":
            v = x(2)
            for k in a[3:]:
                v = v[k]
            return copyv(v)
        if name == "CompositeConstruct":
            out = []
            for k in range(2, len(a)):
                e = x(k)
                if rt[0] == "Vector" and type(e) is list:
                    out.extend(e)
                else:
                    out.append(copyv(e))
            return out
        if name == "CompositeInsert":
            obj, comp = copyv(x(2)), copyv(x(3))
            t = comp
            for k in a[4:-1]:
                t = t[k]
            t[a[-1]] = obj
            return comp
        if name == "VectorShuffle":
            both = list(x(2)) + list(x(3))
            return [both[k] if k != M32 else F(0) for k in a[4:]]
        if name in ("CopyObject", "CopyLogical"): return copyv(x(2))
        if name == "Undef": return m.zero(a[0])
        if name == "Image": return x(2)
        if name == "ImageSampleImplicitLod": return x(2).sample(x(3))  # fragment stage, single-level textures: LOD 0
        if name == "ImageSampleExplicitLod":
            if isinstance(x(2), Texture3DMips):  # operands: mask (a[4], Lod = 0x2), lod id (a[5])
                return x(2).sample(x(3), x(5) if (a[4] & 2) else F(0))
            return x(2).sample(x(3))
        if name == "ImageGather": return x(2).gather(x(3), s32(x(4)))
        if name == "ImageFetch": return x(2).fetch(x(3))
        if name == "ImageQuerySizeLod": return [x(2).W, x(2).H]
        if name == "ImageRead": return x(2).read(x(3))
        if name == "ExtInst": return self.ext(GLSL.get(a[3]), [V(k) for k in a[4:]], a[3])
        raise NotImplementedError(name)

    def ext(self, op, o, num):
        vec = lambda f, *ops: [f(*e) for e in zip(*ops)] if type(ops[0]) is list else f(*ops)
        if op == "FAbs": return vec(lambda p: F(abs(p)), o[0])
        if op == "Floor": return vec(lambda p: F(math.floor(p)), o[0])
        if op == "Fract": return vec(lambda p: F(p - F(math.floor(p))), o[0])
        if op == "Sin": return vec(lambda p: F(math.sin(float(p))), o[0])
        if op == "Cos": return vec(lambda p: F(math.cos(float(p))), o[0])
        if op == "Sqrt": return vec(lambda p: F(np.sqrt(p)), o[0])
        if op == "Pow":
            def pw(p, q):
                try:
                    return F(math.pow(float(p), float(q)))
                except (ValueError, OverflowError):
                    return F(np.nan)
            return vec(pw, o[0], o[1])
        if op == "Step": return vec(lambda edge, p: F(0.0) if p < edge else F(1.0), o[0], o[1])
        if op == "FMin": return vec(gmin, o[0], o[1])
        if op == "FMax": return vec(gmax, o[0], o[1])
        if op == "FClamp": return vec(gclamp, o[0], o[1], o[2])
        if op == "UMin": return vec(lambda p, q: min(p & M32, q & M32), o[0], o[1])
        if op == "UMax": return vec(lambda p, q: max(p & M32, q & M32), o[0], o[1])
        if op == "SMin": return vec(lambda p, q: min(s32(p), s32(q)) & M32, o[0], o[1])
        if op == "SMax": return vec(lambda p, q: max(s32(p), s32(q)) & M32, o[0], o[1])
        if op == "SClamp": return vec(lambda p, lo, hi: min(max(s32(p), s32(lo)), s32(hi)) & M32, o[0], o[1], o[2])
        if op == "UClamp": return vec(lambda p, lo, hi: min(max(p, lo), hi) & M32, o[0], o[1], o[2])
        if op == "FMix": return vec(lambda p, q, t: F(F(p * F(F(1) - t)) + F(q * t)), o[0], o[1], o[2])  # x*(1-a) + y*a, unfused
        if op in ("Length", "Distance", "Normalize"):
            v = o[0] if op != "Distance" else [F(p - q) for p, q in zip(o[0], o[1])]
            acc = F(v[0] * v[0])
            for k in range(1, len(v)):
                acc = F(acc + F(v[k] * v[k]))
            if op != "Normalize":
                return F(np.sqrt(acc))
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = F(np.divide(F(1), F(np.sqrt(acc))))
                return [F(e * inv) for e in v]
        if op == "Cross":  # (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y), every product and difference rounded
            p, q = o[0], o[1]
            return [F(F(p[1] * q[2]) - F(q[1] * p[2])), F(F(p[2] * q[0]) - F(q[2] * p[0])), F(F(p[0] * q[1]) - F(q[0] * p[1]))]
        if op == "Reflect":  # I - 2 * dot(N, I) * N (the factor 2 is exact, so its position does not matter)
            i, n = o[0], o[1]
            d = F(n[0] * i[0])
            for k in range(1, len(i)):
                d = F(d + F(n[k] * i[k]))
            t = F(F(2) * d)
            return [F(e - F(t * nn)) for e, nn in zip(i, n)]
        if op == "MatrixInverse": return inverse4(o[0])
        raise NotImplementedError(f"GLSL.std.450 {num}")


def inverse4(cols):
    """Cofactor expansion of the numerics contract (same operation order as the oracle's inverse4)."""
    A = lambda r, c: cols[c][r]
    mul, sub, add = (lambda p, q: F(p * q)), (lambda p, q: F(p - q)), (lambda p, q: F(p + q))
    d2 = lambda a, b, c, d: sub(mul(a, b), mul(c, d))
    s0 = d2(A(0, 0), A(1, 1), A(1, 0), A(0, 1)); s1 = d2(A(0, 0), A(1, 2), A(1, 0), A(0, 2)); s2 = d2(A(0, 0), A(1, 3), A(1, 0), A(0, 3))
    s3 = d2(A(0, 1), A(1, 2), A(1, 1), A(0, 2)); s4 = d2(A(0, 1), A(1, 3), A(1, 1), A(0, 3)); s5 = d2(A(0, 2), A(1, 3), A(1, 2), A(0, 3))
    c5 = d2(A(2, 2), A(3, 3), A(3, 2), A(2, 3)); c4 = d2(A(2, 1), A(3, 3), A(3, 1), A(2, 3)); c3 = d2(A(2, 1), A(3, 2), A(3, 1), A(2, 2))
    c2 = d2(A(2, 0), A(3, 3), A(3, 0), A(2, 3)); c1 = d2(A(2, 0), A(3, 2), A(3, 0), A(2, 2)); c0 = d2(A(2, 0), A(3, 1), A(3, 0), A(2, 1))
    det = add(sub(add(add(sub(mul(s0, c5), mul(s1, c4)), mul(s2, c3)), mul(s3, c2)), mul(s4, c1)), mul(s5, c0))
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = F(np.divide(F(1), det))
    neg = lambda p: F(-p)
    t3 = lambda a, x, b, y, c, z, sg: mul((add(sub(mul(a, x), mul(b, y)), mul(c, z)) if sg > 0 else sub(add(mul(neg(a), x), mul(b, y)), mul(c, z))), inv)
    B = [[None] * 4 for _ in range(4)]  # B[r][c]
    B[0][0] = t3(A(1, 1), c5, A(1, 2), c4, A(1, 3), c3, +1); B[0][1] = t3(A(0, 1), c5, A(0, 2), c4, A(0, 3), c3, -1)
    B[0][2] = t3(A(3, 1), s5, A(3, 2), s4, A(3, 3), s3, +1); B[0][3] = t3(A(2, 1), s5, A(2, 2), s4, A(2, 3), s3, -1)
    B[1][0] = t3(A(1, 0), c5, A(1, 2), c2, A(1, 3), c1, -1); B[1][1] = t3(A(0, 0), c5, A(0, 2), c2, A(0, 3), c1, +1)
    B[1][2] = t3(A(3, 0), s5, A(3, 2), s2, A(3, 3), s1, -1); B[1][3] = t3(A(2, 0), s5, A(2, 2), s2, A(2, 3), s1, +1)
    B[2][0] = t3(A(1, 0), c4, A(1, 1), c2, A(1, 3), c0, +1); B[2][1] = t3(A(0, 0), c4, A(0, 1), c2, A(0, 3), c0, -1)
    B[2][2] = t3(A(3, 0), s4, A(3, 1), s2, A(3, 3), s0, +1); B[2][3] = t3(A(2, 0), s4, A(2, 1), s2, A(2, 3), s0, -1)
    B[3][0] = t3(A(1, 0), c3, A(1, 1), c1, A(1, 2), c0, -1); B[3][1] = t3(A(0, 0), c3, A(0, 1), c1, A(0, 2), c0, +1)
    B[3][2] = t3(A(3, 0), s3, A(3, 1), s1, A(3, 2), s0, -1); B[3][3] = t3(A(2, 0), s3, A(2, 1), s1, A(2, 2), s0, +1)
    return [[B[r][c] for r in range(4)] for c in range(4)]


def dispatch(mod: Module, groups, on_group=None):
    """Run workgroups `groups` (iterable of (gx, gy, gz)).  Returns total executed SPIR-V instructions."""
    lx, ly, lz = mod.local_size
    total = 0
    for (gx, gy, gz) in groups:
        shared = {}
        invs = []
        for z in range(lz):
            for y in range(ly):
                for x in range(lx):
                    b = {"WorkgroupId": [gx, gy, gz], "LocalInvocationId": [x, y, z], "GlobalInvocationId": [gx * lx + x, gy * ly + y, gz * lz + z],
                         "LocalInvocationIndex": x + y * lx + z * lx * ly, "WorkgroupSize": [lx, ly, lz], "NumWorkgroups": [0, 0, 0]}
                    inv = Invocation(mod, shared, b)
                    invs.append((inv, inv.run()))
        live = invs
        while live:  # lock-step between barriers
            nxt = []
            for inv, gen in live:
                try:
                    next(gen)
                    nxt.append((inv, gen))
                except StopIteration:
                    pass
            live = nxt
        total += sum(inv.count for inv, _ in invs)
        if on_group:
            on_group((gx, gy, gz))
    return total
