"""Normalised C prototypes out of preprocessed headers -- shared by tests/golden/make_golden.py (reference side, needs
/root/reference) and tests/test_abi_and_host.py (our side, include/staple_b200.h).  A prototype is (return type, [parameter
types]) with qualifiers (const, __restrict) and parameter names stripped: what the C ABI sees."""
import re
import subprocess

QUAL = {"const", "__restrict", "restrict", "__restrict__", "struct", "volatile"}
BASE = {"int", "double", "float", "char", "long", "unsigned", "void"}


def norm_param(p):
    p = p.replace("*", " * ").replace("[", " [ ").replace("]", " ] ")
    toks = [t for t in p.split() if t not in QUAL]
    if not toks or toks == ["void"]:
        return None
    if "[" in toks:                       # `type name[]` -> pointer
        i = toks.index("[")
        toks = toks[:i]
        if len(toks) > 1 and re.match(r"^[A-Za-z_]\w*$", toks[-1]):
            toks = toks[:-1]
        return "".join(toks) + "*"
    if len(toks) > 1 and re.match(r"^[A-Za-z_]\w*$", toks[-1]) and toks[-1] not in BASE:
        toks = toks[:-1]
    return "".join(toks)


def prototypes(preprocessed):
    text = re.sub(r"^#.*$", "", preprocessed, flags=re.M)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b([A-Za-z_]\w*)\s*\(([^()]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        rt = "".join(t for t in ret.replace("*", " * ").split() if t not in QUAL | {"extern", "static", "inline"})
        params = [q for q in (norm_param(a) for a in args.split(",")) if q is not None]
        out[name] = [rt, params]
    return out


def preprocess(source_text, flags=()):
    r = subprocess.run(["gcc", "-E", "-P", "-std=gnu99", "-w", "-x", "c", "-"] + list(flags), input=source_text,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-2000:])
    return r.stdout
