"""A small PTX interpreter for ONE warp — test infrastructure: it executes the kernels `fp_codegen.cpp` generates (fp_filter / fp_project)
on the CPU, lane by lane in lock step, so that the SEMANTICS of generated PTX (not only that ptxas assembles it) are checked without a GPU
against python / the oracle.  It knows exactly the instructions the generator emits (tests/test_fp_codegen.py lists them and fails on a new
one); branches must be warp-uniform, which the generated kernels guarantee (unit / round loops).

Registers are numpy arrays of 32 lanes (uint64 bit patterns; predicates as bool).  Global memory is one arena; addresses are arena offsets
plus a base, so a wild address raises instead of reading something."""
import re

import numpy as np

LANES = 32
BASE = 0x10000000
M64 = (1 << 64) - 1


class Arena:
    def __init__(self, size=1 << 22):
        self.mem = np.zeros(size, dtype=np.uint8)
        self.top = 256

    def put(self, array) -> int:
        raw = np.ascontiguousarray(array).view(np.uint8).reshape(-1)
        at = (self.top + 255) & ~255
        assert at + raw.size + 256 <= self.mem.size, "arena too small"
        self.mem[at:at + raw.size] = raw
        self.top = at + raw.size + 64               # a gap between buffers: an overrun lands in zeros, not in the next column
        return BASE + at

    def alloc(self, nbytes) -> int:
        return self.put(np.zeros(max(nbytes, 1), dtype=np.uint8))

    def read(self, address, nbytes, dtype):
        at = address - BASE
        assert 0 <= at and at + nbytes <= self.mem.size, f"load outside the arena: {address:#x}"
        return self.mem[at:at + nbytes].view(dtype)

    def load(self, address, width) -> int:
        at = address - BASE
        assert 0 <= at and at + width <= self.mem.size, f"load outside the arena: {address:#x}"
        return int.from_bytes(self.mem[at:at + width].tobytes(), "little")

    def store(self, address, width, value):
        at = address - BASE
        assert 0 <= at and at + width <= self.mem.size, f"store outside the arena: {address:#x}"
        self.mem[at:at + width] = np.frombuffer(int(value & ((1 << (8 * width)) - 1)).to_bytes(width, "little"), dtype=np.uint8)


def _f64(bits):
    return np.asarray(bits, dtype=np.uint64).view(np.float64)


def _bits(f):
    return np.asarray(f, dtype=np.float64).view(np.uint64)


def _signed(x, width):
    x = np.asarray(x, dtype=np.uint64)
    if width == 64:
        return x.view(np.int64)
    return (x & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32).astype(np.int64)


class Kernel:
    def __init__(self, ptx: str, name: str):
        m = re.search(r"\.visible \.entry " + name + r"\((.*?)\)\s*\{(.*?)\n\}", ptx, re.S)
        assert m, f"kernel {name} not found"
        self.params = re.findall(r"\.param \.u64 (\w+)", m.group(1))
        self.code, self.labels = [], {}
        for raw in m.group(2).splitlines():
            line = raw.split("//")[0].strip()
            if not line or line.startswith(".reg"):
                continue
            if line.endswith(":"):
                self.labels[line[:-1]] = len(self.code)
                continue
            assert line.endswith(";"), line
            line = line[:-1].strip()
            guard = None
            g = re.match(r"@(!?)(%\w+)\s+(.*)", line)
            if g:
                guard, line = (g.group(2), g.group(1) == "!"), g.group(3)
            op, _, rest = line.partition(" ")
            args = [a.strip() for a in re.split(r",(?![^\[]*\])", rest)] if rest else []
            self.code.append((guard, op, args))
        self.opcodes = {op for _, op, _ in self.code}

    def run(self, arena: Arena, params: dict):
        assert set(params) == set(self.params), (sorted(params), self.params)
        reg = {}
        lane = np.arange(LANES, dtype=np.uint64)
        special = {"%tid.x": lane, "%ntid.x": np.full(LANES, LANES, dtype=np.uint64), "%ctaid.x": np.zeros(LANES, dtype=np.uint64),
                   "%nctaid.x": np.ones(LANES, dtype=np.uint64), "%lanemask_lt": (np.uint64(1) << lane) - np.uint64(1)}

        def val(a):
            if a in special:
                return special[a]
            if a.startswith("%"):
                return reg[a]
            if a.startswith("0x"):
                return np.full(LANES, int(a, 16) & M64, dtype=np.uint64)
            if re.fullmatch(r"-?\d+", a):
                return np.full(LANES, int(a) & M64, dtype=np.uint64)
            raise AssertionError(f"operand {a}")

        def address(a):
            m = re.fullmatch(r"\[(%\w+)(?:\+(\d+))?\]", a)
            assert m, a
            return val(m.group(1)) + np.uint64(int(m.group(2) or 0))

        def write(dst, value, mask, pred=False):
            value = np.asarray(value)
            if dst not in reg:
                reg[dst] = np.zeros(LANES, dtype=bool if pred else np.uint64)
            reg[dst] = np.where(mask, value.astype(bool) if pred else value.astype(np.uint64), reg[dst])

        pc, steps = 0, 0
        while True:
            steps += 1
            assert steps < 5_000_000, "runaway kernel"
            guard, op, a = self.code[pc]
            pc += 1
            mask = np.ones(LANES, dtype=bool)
            if guard:
                mask = reg[guard[0]] ^ guard[1] if guard[1] else reg[guard[0]].copy()
            parts = op.split(".")
            base = parts[0]
            ty = parts[-1]
            width = 64 if ty.endswith("64") else (8 if ty.endswith("8") else 32)
            wmask = np.uint64((1 << width) - 1) if width < 64 else np.uint64(M64)
            if base == "ret;" or op == "ret":
                return
            if base == "bra":
                assert mask.all() or not mask.any(), "divergent branch"
                if mask.all():
                    pc = self.labels[a[0]]
                continue
            if base == "ld" and parts[1] == "param":
                write(a[0], np.full(LANES, params[a[1].strip("[]")] & M64, dtype=np.uint64), mask)
            elif base == "cvta":
                write(a[0], val(a[1]), mask)
            elif base == "mov":
                write(a[0], val(a[1]) if ty != "pred" else reg[a[1]], mask, pred=ty == "pred")
            elif base in ("add", "sub", "mul", "div") and ty == "f64":
                x, y = _f64(val(a[1])), _f64(val(a[2]))
                with np.errstate(all="ignore"):
                    r = x + y if base == "add" else x - y if base == "sub" else x * y if base == "mul" else x / y
                write(a[0], _bits(r), mask)
            elif base in ("add", "sub"):
                x, y = val(a[1]), val(a[2])
                write(a[0], ((x + y) if base == "add" else (x - y)) & wmask, mask)
            elif base == "mul" and parts[1] == "lo":
                write(a[0], (val(a[1]) * val(a[2])) & wmask, mask)
            elif base == "mul" and parts[1] == "hi":
                x, y = _signed(val(a[1]), 64), _signed(val(a[2]), 64)
                write(a[0], np.array([((int(p) * int(q)) >> 64) & M64 for p, q in zip(x, y)], dtype=np.uint64), mask)
            elif base == "mad":
                write(a[0], (val(a[1]) * val(a[2]) + val(a[3])) & wmask, mask)
            elif base == "div":       # s64
                x, y = _signed(val(a[1]), 64), _signed(val(a[2]), 64)
                write(a[0], np.array([(abs(int(p)) // abs(int(q)) * (1 if (p < 0) == (q < 0) else -1)) & M64 if q else 0 for p, q in zip(x, y)], dtype=np.uint64), mask)
            elif base == "neg":
                write(a[0], (np.uint64(0) - val(a[1])) & wmask, mask)
            elif base == "min":
                write(a[0], np.minimum(val(a[1]), val(a[2])), mask)
            elif base in ("and", "or", "xor") and ty == "pred":
                x, y = reg[a[1]], reg[a[2]]
                write(a[0], (x & y) if base == "and" else (x | y) if base == "or" else (x ^ y), mask, pred=True)
            elif base == "not" and ty == "pred":
                write(a[0], ~reg[a[1]], mask, pred=True)
            elif base in ("and", "or", "xor"):
                x, y = val(a[1]), val(a[2])
                write(a[0], ((x & y) if base == "and" else (x | y) if base == "or" else (x ^ y)) & wmask, mask)
            elif base == "shl":
                write(a[0], (val(a[1]) << (val(a[2]) & np.uint64(63))) & wmask, mask)
            elif base == "shr":
                sh = val(a[2]) & np.uint64(63)
                if ty.startswith("s"):
                    write(a[0], (_signed(val(a[1]), width) >> sh.astype(np.int64)).astype(np.int64).view(np.uint64) & wmask, mask)
                else:
                    write(a[0], ((val(a[1]) & wmask) >> sh), mask)
            elif base == "popc":
                write(a[0], np.array([bin(int(v) & 0xFFFFFFFF).count("1") for v in val(a[1])], dtype=np.uint64), mask)
            elif base == "cvt":
                src = val(a[1])
                if op == "cvt.s64.s32":
                    write(a[0], _signed(src, 32).view(np.uint64), mask)
                elif op in ("cvt.u64.u32", "cvt.u32.u64"):
                    write(a[0], src & np.uint64(0xFFFFFFFF), mask)
                elif op == "cvt.rn.f64.s64":
                    write(a[0], _bits(_signed(src, 64).astype(np.float64)), mask)
                else:
                    raise AssertionError(op)
            elif base == "setp":
                cmp, x, y = parts[1], val(a[1]), val(a[2])
                if ty == "f64":
                    p, q = _f64(x), _f64(y)
                    with np.errstate(all="ignore"):
                        r = {"eq": p == q, "ne": (p != q) & ~np.isnan(p) & ~np.isnan(q), "neu": ~(p == q), "lt": p < q, "le": p <= q, "gt": p > q, "ge": p >= q,
                             "nan": np.isnan(p) | np.isnan(q)}[cmp]
                elif ty[0] == "s":
                    p, q = _signed(x, width), _signed(y, width)
                    r = {"eq": p == q, "ne": p != q, "lt": p < q, "le": p <= q, "gt": p > q, "ge": p >= q}[cmp]
                else:
                    p, q = x & wmask, y & wmask
                    r = {"eq": p == q, "ne": p != q, "lt": p < q, "le": p <= q, "gt": p > q, "ge": p >= q}[cmp]
                write(a[0], r, mask, pred=True)
            elif base == "selp":
                write(a[0], np.where(reg[a[3]], val(a[1]), val(a[2])) & wmask, mask)
            elif base == "vote":
                bits = sum(1 << i for i in range(LANES) if reg[a[1]][i])
                write(a[0], np.full(LANES, bits, dtype=np.uint64), mask)
            elif base == "ld" and parts[1] == "global":
                nbytes = width // 8
                adr = address(a[1])
                out = reg.get(a[0], np.zeros(LANES, dtype=np.uint64)).copy()
                for i in range(LANES):
                    if mask[i]:
                        out[i] = arena.load(int(adr[i]), nbytes)
                reg[a[0]] = out
            elif base == "st":
                nbytes = width // 8
                adr, v = address(a[0]), val(a[1])
                for i in range(LANES):
                    if mask[i]:
                        arena.store(int(adr[i]), nbytes, int(v[i]))
            elif base == "red":       # red.global.or.b32 [addr], imm
                adr, v = address(a[0]), val(a[1])
                for i in range(LANES):
                    if mask[i]:
                        arena.store(int(adr[i]), 4, arena.load(int(adr[i]), 4) | int(v[i]))
            else:
                raise AssertionError(f"instruction not interpreted: {op}")
