"""The inline-PTX carry chains of nova_b200/csrc/field.cuh, INTERPRETED on the CPU.

tests/test_host_field.py validates the `#else` (host emulation) branch of every chain against Python integers; the
device executes the `#ifdef __CUDA_ARCH__` branch.  This test closes the gap between the two without a GPU: it
extracts each `asm(...)` statement from the source, parses its template and operand lists, interprets the PTX
(mad.lo/hi(.cc), madc, add(.cc), addc, sub(.cc), subc on 32-bit registers with the carry flag) and checks the
result against the arithmetic the chain is specified to perform.  A wrong operand index, a swapped lo/hi or a
missing carry would show up here.  (What remains GPU-only is ptxas and the hardware.)"""
import os
import random
import re

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "nova_b200", "csrc", "field.cuh")
M32 = (1 << 32) - 1


# ------------------------------------------------------------------------------------------------
# extraction: asm("..." "..." : outputs : inputs);
# ------------------------------------------------------------------------------------------------
def _asm_statements(text):
    out, i = [], 0
    while True:
        i = text.find("asm(", i)
        if i < 0:
            return out
        depth, j, in_str = 0, i + 3, False
        while True:  # matching parenthesis, skipping string literals
            ch = text[j]
            if in_str:
                if ch == "\\":
                    j += 1
                elif ch == '"':
                    in_str = False
            elif ch == '"':
                in_str = True
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        out.append((text.count("\n", 0, i) + 1, text[i + 4:j]))
        i = j


def _split_top(s, sep):
    parts, depth, cur, in_str = [], 0, "", False
    k = 0
    while k < len(s):
        ch = s[k]
        if in_str:
            cur += ch
            if ch == "\\":
                k += 1
                cur += s[k]
            elif ch == '"':
                in_str = False
        elif ch == '"':
            in_str = True
            cur += ch
        elif ch in "([":
            depth += 1
            cur += ch
        elif ch in ")]":
            depth -= 1
            cur += ch
        elif ch == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
        k += 1
    parts.append(cur)
    return parts


def parse_asm(body):
    """-> (instructions, outputs [(constraint, expr)], inputs [(constraint, expr)])"""
    secs = _split_top(body, ":")
    template = "".join(bytes(m, "utf-8").decode("unicode_escape") for m in re.findall(r'"((?:[^"\\]|\\.)*)"', secs[0]))
    ops = []
    for sec in secs[1:3]:
        lst = []
        for item in _split_top(sec, ","):
            item = item.strip()
            if not item:
                continue
            m = re.match(r'"([^"]+)"\s*\((.*)\)$', item, re.S)
            lst.append((m.group(1), m.group(2).strip()))
        ops.append(lst)
    while len(ops) < 2:
        ops.append([])
    instrs = []
    for line in re.split(r"[;\n]", template.replace("{", "\n").replace("}", "\n")):
        line = line.strip()
        if line and not line.startswith(".reg"):
            instrs.append(line)
    return instrs, ops[0], ops[1]


# ------------------------------------------------------------------------------------------------
# interpretation
# ------------------------------------------------------------------------------------------------
def run_ptx(instrs, values):
    """values: list indexed by operand number (outputs first, then inputs).  Returns the updated list."""
    regs = {}
    vals = list(values)
    cf = 0

    def get(tok):
        tok = tok.strip()
        if tok.startswith("%"):
            return vals[int(tok[1:])]
        if re.fullmatch(r"-?\d+|0x[0-9a-fA-F]+", tok):
            return int(tok, 0) & M32
        return regs[tok]

    def put(tok, v):
        tok = tok.strip()
        if tok.startswith("%"):
            vals[int(tok[1:])] = v & M32
        else:
            regs[tok] = v & M32
    for ins in instrs:
        op, rest = ins.split(None, 1)
        args = [a.strip() for a in rest.split(",")]
        parts = op.split(".")
        name = parts[0]
        sets_cc = "cc" in parts
        if name in ("mad", "madc"):
            prod = get(args[1]) * get(args[2])
            term = (prod & M32) if "lo" in parts else (prod >> 32)
            t = term + get(args[3]) + (cf if name == "madc" else 0)
        elif name in ("add", "addc"):
            t = get(args[1]) + get(args[2]) + (cf if name == "addc" else 0)
        elif name in ("sub", "subc"):
            t = get(args[1]) - get(args[2]) - (cf if name == "subc" else 0)
        else:
            raise AssertionError(f"unsupported PTX instruction {ins!r}")
        put(args[0], t)
        if sets_cc:
            cf = (1 if t < 0 else 0) if name.startswith("sub") else (t >> 32) & 1
        assert "u32" in parts, ins
    return vals


def bind(ops_out, ops_in, env):
    """Evaluate the C operand expressions (valid Python once template parameters are in `env`)."""
    return [eval(e, {}, env) for _, e in ops_out + ops_in]


def write_back(ops_out, vals, env):
    for (_, expr), v in zip(ops_out, vals):
        m = re.match(r"(\w+)\[(.*)\]$", expr)
        if m:
            env[m.group(1)][eval(m.group(2), {}, env)] = v
        else:
            env[expr] = v


@pytest.fixture(scope="module")
def blocks():
    text = open(SRC).read()
    stmts = [(ln, parse_asm(body)) for ln, body in _asm_statements(text)]
    assert len(stmts) == 12, "a chain was added or removed in field.cuh: extend this test"
    return stmts


def limbs(x, n):
    return [(x >> (32 * i)) & M32 for i in range(n)]


def value(ls):
    return sum(v << (32 * i) for i, v in enumerate(ls))


RND = random.Random(20240917)
EDGE = [0, 1, M32, M32 - 1, 1 << 31]


def r32():
    return RND.choice(EDGE) if RND.random() < 0.3 else RND.getrandbits(32)


def test_chain_mad_with_and_without_carry_in(blocks):
    """X[OFF .. OFF+8] += (x0, x1, x2, x3) * y, product k on limbs (OFF+2k, OFF+2k+1), carry-in = carry32(ca + cb)."""
    for idx, cin in ((0, True), (1, False)):
        instrs, outs, ins = blocks[idx][1]
        assert len(outs) == 9 and len(ins) == (7 if cin else 5)
        for _ in range(400):
            OFF = RND.randrange(0, 8)
            X = [r32() for _ in range(17)]
            X[OFF + 8] = RND.randrange(0, 8)  # the small carry counter the multiplier keeps there
            env = dict(OFF=OFF, X=list(X), x0=r32(), x1=r32(), x2=r32(), x3=r32(), y=r32(), ca=r32(), cb=r32())
            vals = run_ptx(instrs, bind(outs, ins, env))
            write_back(outs, vals, env)
            cin_v = ((env["ca"] + env["cb"]) >> 32) if cin else 0
            exp = value(X[OFF:OFF + 9]) + sum(env[f"x{k}"] * env["y"] << (64 * k) for k in range(4)) + cin_v
            assert value(env["X"][OFF:OFF + 9]) == exp and exp < 1 << 288
            assert env["X"][:OFF] == X[:OFF] and env["X"][OFF + 9:] == X[OFF + 9:]


def test_add8_add8_cin_sub8(blocks):
    instrs, outs, ins = blocks[2][1]  # add8: r = a + b (carry-out dropped)
    for _ in range(300):
        env = dict(r=[0] * 8, a=[r32() for _ in range(8)], b=[r32() for _ in range(8)])
        write_back(outs, run_ptx(instrs, bind(outs, ins, env)), env)
        assert value(env["r"]) == (value(env["a"]) + value(env["b"])) % (1 << 256)
    instrs, outs, ins = blocks[3][1]  # add8_cin: r = a + b + carry32(ca + cb)
    for _ in range(300):
        env = dict(r=[0] * 8, a=[r32() for _ in range(8)], b=[r32() for _ in range(8)], ca=r32(), cb=r32())
        write_back(outs, run_ptx(instrs, bind(outs, ins, env)), env)
        assert value(env["r"]) == (value(env["a"]) + value(env["b"]) + ((env["ca"] + env["cb"]) >> 32)) % (1 << 256)
    instrs, outs, ins = blocks[4][1]  # sub8: r = a - b, borrow = 0xffffffff on borrow
    for _ in range(300):
        env = dict(r=[0] * 8, a=[r32() for _ in range(8)], b=[r32() for _ in range(8)], borrow=0)
        if RND.random() < 0.1:
            env["b"] = list(env["a"])
        write_back(outs, run_ptx(instrs, bind(outs, ins, env)), env)
        d = value(env["a"]) - value(env["b"])
        assert value(env["r"]) == d % (1 << 256) and env["borrow"] == (M32 if d < 0 else 0)


def test_carry_save_primitives(blocks):
    instrs, outs, ins = blocks[5][1]  # mad_cs: (hi:lo) += x * y, k += carry-out
    for _ in range(300):
        env = dict(lo=r32(), hi=r32(), k=RND.randrange(0, 100), x=r32(), y=r32())
        before = dict(env)
        write_back(outs, run_ptx(instrs, bind(outs, ins, env)), env)
        t = (before["hi"] << 32 | before["lo"]) + before["x"] * before["y"]
        assert (env["hi"] << 32 | env["lo"]) == t % (1 << 64) and env["k"] == before["k"] + (t >> 64)
    instrs, outs, ins = blocks[6][1]  # retire_cs: k_next += carry32(e + o) + kz
    for _ in range(300):
        env = dict(k_next=RND.randrange(0, 100), t=0, e=r32(), o=r32(), kz=RND.randrange(0, 2))
        before = dict(env)
        write_back(outs, run_ptx(instrs, bind(outs, ins, env)), env)
        assert env["k_next"] == before["k_next"] + ((before["e"] + before["o"]) >> 32) + before["kz"]


def test_variable_length_chains_of_the_dedicated_squaring(blocks):
    """chain_mad_n<OFF, NP>: X[OFF .. OFF+2NP-1] += (x0 .. x_{NP-1}) * y, carry into X[OFF+2NP]; chain_sqr_diag:
    X[0..15] += a_k^2 on limbs (2k, 2k+1), carry into X[16]."""
    for NP in (1, 2, 3, 4):
        instrs, outs, ins = blocks[6 + NP][1]
        assert len(outs) == 2 * NP + 1 and len(ins) == NP + 1
        for _ in range(300):
            OFF = RND.randrange(0, 17 - 2 * NP)
            X = [r32() for _ in range(17)]
            X[OFF + 2 * NP] = RND.randrange(0, 4)  # untouched / small when the chain is issued (see field.cuh)
            env = dict(OFF=OFF, X=list(X), x0=r32(), x1=r32(), x2=r32(), x3=r32(), y=r32())
            write_back(outs, run_ptx(instrs, bind(outs, ins, env)), env)
            exp = value(X[OFF:OFF + 2 * NP + 1]) + sum(env[f"x{k}"] * env["y"] << (64 * k) for k in range(NP))
            assert value(env["X"][OFF:OFF + 2 * NP + 1]) == exp
            assert env["X"][:OFF] == X[:OFF] and env["X"][OFF + 2 * NP + 1:] == X[OFF + 2 * NP + 1:]
    instrs, outs, ins = blocks[11][1]
    assert len(outs) == 17 and len(ins) == 8

    class A:  # the operand expressions read a.l[k]
        pass
    for _ in range(300):
        a = A()
        a.l = [r32() for _ in range(8)]
        X = [r32() for _ in range(16)] + [RND.randrange(0, 3)]
        env = dict(X=list(X), a=a)
        write_back(outs, run_ptx(instrs, bind(outs, ins, env)), env)
        assert value(env["X"]) == value(X) + sum(a.l[k] * a.l[k] << (64 * k) for k in range(8))
