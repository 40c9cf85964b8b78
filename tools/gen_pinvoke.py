"""C# P/Invoke declarations generated from include/yolob200.h (the binding a YoloSharp maintainer would add next to
YoloSharp/Utils).  python tools/gen_pinvoke.py [name ...] prints the declarations of the named
functions (default: all)."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declarations():
    h = open(os.path.join(ROOT, "include", "yolob200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    out = []
    for m in re.finditer(r"^\s*((?:const\s+)?[A-Za-z_][A-Za-z0-9_]*\s*\**)\s*(yb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h, flags=re.M):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        out.append((ret, name, [] if args in ("", "void") else [a.strip() for a in args.split(",")]))
    return out


def cs_type(c, is_ret=False):
    c = c.strip()
    if "*" in c:
        if re.match(r"const\s+char\s*\*$", c) and not is_ret:
            return "string"
        return "IntPtr"
    return {"int32_t": "int", "int64_t": "long", "uint8_t": "byte", "float": "float", "double": "double", "void": "void", "int": "int",
            "uint32_t": "uint", "uint64_t": "ulong", "size_t": "UIntPtr"}.get(c, "IntPtr")


def camel(n):
    parts = n.split("_")
    return parts[0] + "".join(p[:1].upper() + p[1:] for p in parts[1:])


def render(ret, name, args):
    ps = []
    for a in args:
        arr = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)\[\d*\]$", a)  # `int32_t chw[3]`: an array parameter is a pointer
        if arr:
            ps.append(f"int[] {camel(arr.group(2))}" if "int32_t" in arr.group(1) else f"IntPtr {camel(arr.group(2))}")
            continue
        m = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)$", a)
        ps.append(f"{cs_type(m.group(1))} {camel(m.group(2))}")
    cs_kw = {"out", "in", "ref", "params", "base", "lock", "event", "string", "object"}
    ps = [p if p.split()[-1] not in cs_kw else p.rsplit(" ", 1)[0] + " @" + p.split()[-1] for p in ps]
    return f"[DllImport(Lib)] internal static extern {cs_type(ret, True)} {name}({', '.join(ps)});"


if __name__ == "__main__":
    want = set(sys.argv[1:])
    for ret, name, args in declarations():
        if not want or name in want:
            print(render(ret, name, args))
