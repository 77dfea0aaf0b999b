"""TEST INFRASTRUCTURE.  Translates the product's CUDA sources (huff_encoding_b200/csrc) into C++ that g++ compiles against
tests/emu/include (the CPU execution model).  The product sources are never edited for this; what cannot be compiled for
the host is rewritten here, mechanically and loudly (anything unexpected is an error, not a guess):

  kernel<<<grid, block, smem, stream>>>(args);   -> hb_emu::launch(grid, block, smem, "kernel", [=] { kernel(args); });
  extern __shared__ ... NAME[];                  -> NAME = the dynamic part of the CTA's shared-memory arena
  __shared__ T NAME[...];                        -> a reference into the static part of the arena
  helper functions written in PTX                -> the C++ bodies in PTX_HELPERS below (same semantics, plus the
                                                    alignment / bounds rules the PTX instruction imposes, checked)
  the few inline PTX statements inside kernels   -> the C++ in PTX_STATEMENTS
"""
from __future__ import annotations

import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "huff_encoding_b200", "csrc")

_SM = "hb_emu::smem_base()"

# function name -> C++ body (parameter names are those of the product's signature)
PTX_HELPERS = {
    # hb_common.cuh
    "ld_stream_u4": 'if (reinterpret_cast<uintptr_t>(p) & 15) hb_emu::trap("ld.global.v4.u32: address not 16-byte aligned"); return *p;',
    "ld_stream_u32": "return *p;",
    "ld_stream_256": 'if (reinterpret_cast<uintptr_t>(p) & 31) hb_emu::trap("ld.global.v8.b32: address not 32-byte aligned"); '
                     "u32x8 r; memcpy(&r, p, 32); return r;",
    "st_stream_256": 'if (reinterpret_cast<uintptr_t>(p) & 31) hb_emu::trap("st.global.v8.b32: address not 32-byte aligned"); memcpy(p, &r, 32);',
    "st_stream_u32": "*p = v;",
    "st_stream_u4": 'if (reinterpret_cast<uintptr_t>(p) & 15) hb_emu::trap("st.global.v4.u32: address not 16-byte aligned"); *p = v;',
    "ld_acquire_u64": "return __atomic_load_n(p, __ATOMIC_ACQUIRE);",
    "st_release_u64": "__atomic_store_n(p, v, __ATOMIC_RELEASE);",
    "ld_relaxed_u32": "return __atomic_load_n(p, __ATOMIC_RELAXED);",
    # hb_decode.cuh
    "lds32": f'hb_emu::smem_check(a, 4, "ld.shared.u32"); uint32_t v; memcpy(&v, {_SM} + a, 4); return v;',
    "lds16": f'hb_emu::smem_check(a, 2, "ld.shared.u16"); uint16_t v; memcpy(&v, {_SM} + a, 2); return v;',
    "lds8": f'hb_emu::smem_check(a, 1, "ld.shared.u8"); return {_SM}[a];',
    "stg256": 'if (reinterpret_cast<uintptr_t>(p) & 31) hb_emu::trap("st.global.v8.b32: address not 32-byte aligned"); memcpy(p, v, 32);',
    # hb_encode.cuh
    "lds_u32": f'hb_emu::smem_check(a, 4, "ld.shared.u32"); uint32_t r; memcpy(&r, {_SM} + a, 4); return r;',
    # hb_hist.cuh
    "hist_red": f'hb_emu::smem_check(addr, 4, "red.shared.add.u32"); uint32_t v; memcpy(&v, {_SM} + addr, 4); v += 1u; memcpy({_SM} + addr, &v, 4);',
    # hb_decode_fused.cuh
    "team_sync": "hb_emu::barrier(team + 1, kFTeam);",
    "sts32": f'hb_emu::smem_check(a, 4, "st.shared.u32"); memcpy({_SM} + a, &v, 4);',
    "ld_relaxed_ull": "return __atomic_load_n(p, __ATOMIC_RELAXED);",
    "st_relaxed_ull": "__atomic_store_n(p, v, __ATOMIC_RELAXED);",
    "refill": "const uint32_t r = (qn ^ q) & 32u; if (r) { w0 = w1; w1 = w2; w2 = lds32(wa); wa += 4; } q = qn;",
    "step_fma": "const uint32_t qn = q + bits; const uint32_t r = (qn ^ q) & 32u; "
                "if (r) { w0 = w1 * k1; w1 = w2 * k1; w2 = lds32(wa); wa = k4 * k1 + wa; } q = qn;",
    "emit_off": "const uint32_t y = x & ~((1u << (32 - EB)) - 1u); return y >> (30 - EB);",
    # mbarrier + bulk copy: hb_emu.cpp (eager or lazy landing of the bytes)
    "mbar_init": "(void)count; hb_emu::mbar_init(bar);",
    "mbar_expect_tx": 'hb_emu::smem_check(bar, 8, "mbarrier.arrive.expect_tx"); (void)bytes;',
    "mbar_wait": "hb_emu::mbar_wait(bar, parity);",
    "bulk_g2s": "hb_emu::bulk_copy(dst, src, bytes, bar);",
}

# exact inline statements (whitespace-normalised) -> C++
PTX_STATEMENTS = [
    (r'asm volatile\("trap;"\);', 'hb_emu::trap("trap instruction reached");'),
    (r'asm volatile\("mov\.u32 %0, %0;" : "\+r"\(b\)\);', "(void)0;"),
    (r'asm volatile\("fence\.mbarrier_init\.release\.cluster;" ::: "memory"\);', "(void)0;"),
    (r'asm volatile\("fence\.proxy\.async\.shared::cta;" ::: "memory"\);', "(void)0;"),
    (r'asm volatile\("prefetch\.global\.L2 \[%0\];" :: "l"\((?P<a>[^;]*)\)\);', "(void)(\\g<a>);"),
    # y = X & MASK
    (r'asm\("and\.b32 %0, %1, %2;" : "=r"\(y\) : "r"\((?P<x>.*?)\), "n"\((?P<m>~\(\(1u << \(32 - kLutBits\)\) - 1u\))\)\);',
     "y = (\\g<x>) & (\\g<m>);"),
    # the ring append of the fused decoder: predicated store of a completed word, accumulator takes the spill-over
    (r'asm volatile\("\{\\n\\t\.reg \.pred f;\\n\\tsetp\.ne\.u32 f, %3, 0;\\n\\t@f st\.shared\.u32 \[%2\], %0;\\n\\t"\s*'
     r'"@f mad\.lo\.u32 %0, %1, %4, 0;\\n\\t\}"\s*: "\+r"\(acc\) : "r"\(hi\), "r"\(wa_ring\), "r"\(fl\), "r"\(k1\) : "memory"\);',
     "if (fl) { sts32(wa_ring, acc); acc = hi * k1; }"),
    (r'asm\("mad\.lo\.u32 %0, %1, %2, %0;" : "\+r"\(wp\) : "r"\(fl\), "r"\(k1\)\);', "wp += fl * k1;"),
]


def _match_brace(s: str, open_idx: int, op: str = "{", cl: str = "}") -> int:
    """index of the bracket closing the one at open_idx (no string/comment awareness needed for these sources' bodies,
    except string literals inside asm, which contain braces: skip over string literals)."""
    depth, i, n = 0, open_idx, len(s)
    while i < n:
        c = s[i]
        if c == '"':
            i += 1
            while s[i] != '"':
                i += 2 if s[i] == "\\" else 1
        elif c == "'" and op != "<":
            i += 1
            while s[i] != "'":
                i += 2 if s[i] == "\\" else 1
        elif c == "/" and s[i + 1] == "/":
            i = s.index("\n", i)
            continue
        elif c == op:
            depth += 1
        elif c == cl:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced " + op)


def replace_helper_bodies(src: str, path: str, used: set) -> str:
    for name, body in PTX_HELPERS.items():
        # a definition: "name(" ... ")" [const] "{" with asm inside the braces
        for m in list(re.finditer(r"\b%s\s*\(" % re.escape(name), src)):
            close = _match_brace(src, m.end() - 1, "(", ")")
            k = close + 1
            rest = src[k:k + 40].lstrip()
            if rest.startswith("const"):
                rest = rest[5:].lstrip()
            if not rest.startswith("{"):
                continue                                   # a call, not a definition
            open_idx = src.index("{", close)
            end = _match_brace(src, open_idx)
            if "asm" not in src[open_idx:end]:
                continue
            src = src[:open_idx] + "{ " + body + " }" + src[end + 1:]
            used.add(name)
            break
    return src


def replace_statements(src: str) -> str:
    for pat, repl in PTX_STATEMENTS:
        src = re.sub(pat, repl, src, flags=re.S)
    return src


_site = [0]


def replace_shared(src: str) -> str:
    # dynamic: extern __shared__ __align__(N) uint8_t NAME[];
    def dyn(m):
        return f"#define {m.group(1)} (hb_emu::dyn_smem())"
    src = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?uint8_t\s+(\w+)\[\];", dyn, src)

    # static: __shared__ TYPE NAME[dims]...;   (TYPE may be two words, e.g. unsigned long long)
    def stat(m):
        typ, name, dims = m.group(1).strip(), m.group(2), m.group(3) or ""
        _site[0] += 1
        t = f"hb_emu_t_{name}_{_site[0]}"
        return (f"typedef {typ} {t}{dims}; {t} &{name} = "
                f"*reinterpret_cast<{t} *>(hb_emu::static_smem(sizeof({t}), alignof({t}), {_site[0]}));")
    src = re.sub(r"(?<!extern )__shared__\s+((?:unsigned\s+long\s+long|\w+))\s+(\w+)((?:\[[^\]]*\])*)\s*;", stat, src)
    if "__shared__" in re.sub(r"//.*", "", src):
        raise ValueError("an unhandled __shared__ declaration remains")
    return src


def replace_launches(src: str) -> str:
    out, i = [], 0
    while True:
        j = src.find("<<<", i)
        if j < 0:
            out.append(src[i:])
            break
        # kernel expression: identifiers, ::, and balanced <...> going backwards from j
        k = j
        depth = 0
        while k > 0:
            c = src[k - 1]
            if c == ">":
                depth += 1
            elif c == "<":
                depth -= 1
            elif depth == 0 and not (c.isalnum() or c in "_:"):
                break
            k -= 1
        kernel = src[k:j]
        cfg_end = src.index(">>>", j)
        cfg = src[j + 3:cfg_end]
        parts, depth, cur = [], 0, ""
        for c in cfg:
            if c in "([":
                depth += 1
            elif c in ")]":
                depth -= 1
            if c == "," and depth == 0:
                parts.append(cur.strip())
                cur = ""
            else:
                cur += c
        parts.append(cur.strip())
        if len(parts) not in (2, 3, 4):
            raise ValueError(f"launch configuration not understood: {cfg}")
        grid, block = parts[0], parts[1]
        smem = parts[2] if len(parts) > 2 else "0"
        a0 = src.index("(", cfg_end)
        if src[cfg_end + 3:a0].strip():
            raise ValueError("launch without an argument list")
        a1 = _match_brace(src, a0, "(", ")")
        args = src[a0 + 1:a1]
        semi = src.index(";", a1)
        out.append(src[i:k])
        out.append(f"hb_emu::launch(hb_emu::Dim3({grid}), hb_emu::Dim3({block}), ({smem}), \"{kernel}\", "
                   f"[=]() {{ {kernel}({args}); }});")
        i = semi + 1
    return "".join(out)


def translate(text: str, path: str, used: set) -> str:
    text = replace_helper_bodies(text, path, used)
    text = replace_statements(text)
    # (a macro named __noinline__ would break libstdc++'s own __attribute__((__noinline__)))
    text = re.sub(r"\b__noinline__\b", "__attribute__((noinline))", text)
    text = replace_shared(text)
    text = replace_launches(text)
    code = re.sub(r"//.*", "", text)
    if re.search(r"\basm\b", code):
        line = code[:re.search(r"\basm\b", code).start()].count("\n") + 1
        raise ValueError(f"{path}:{line}: an inline PTX statement has no C++ counterpart in tests/emu/translate.py")
    return "// GENERATED by tests/emu/translate.py from " + os.path.relpath(path, ROOT) + " -- do not edit\n" + text


def translate_tree(out_dir: str) -> list[str]:
    os.makedirs(out_dir, exist_ok=True)
    used: set = set()
    written = []
    for name in sorted(os.listdir(CSRC)):
        if not name.endswith((".cu", ".cuh", ".cpp", ".h")):
            continue
        path = os.path.join(CSRC, name)
        text = open(path).read()
        if name.endswith((".cu", ".cuh")):
            text = translate(text, path, used)
            text = text.replace('#include "../../include/huffb200.h"', '#include "huffb200.h"')
            if name == "hb_api.cu":
                # the model's ranks are threads of one process: NCCL is tests/emu/nccl_emu.cpp, found next to the model
                if text.count('"libnccl.so.2"') != 2 or text.count('"libnccl.so"') != 1:
                    raise ValueError("hb_api.cu: the dlopen calls for NCCL are not where translate.py expects them")
                emu = os.path.join(out_dir, "libnccl_emu.so")
                text = text.replace('"libnccl.so.2"', f'"{emu}"').replace('"libnccl.so"', f'"{emu}"')
        else:
            text = text.replace('#include "../../include/huffb200.h"', '#include "huffb200.h"')
        dst = os.path.join(out_dir, name[:-3] + ".cpp" if name.endswith(".cu") else name)
        if not os.path.exists(dst) or open(dst).read() != text:
            open(dst, "w").write(text)
        written.append(dst)
    missing = set(PTX_HELPERS) - used
    if missing:
        raise ValueError(f"PTX helper(s) not found in the sources (renamed?): {sorted(missing)}")
    return written


if __name__ == "__main__":
    for p in translate_tree(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_build")):
        print(p)
