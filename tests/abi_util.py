"""Parse include/fdn_b200.h into ctypes-style signature codes (used by the ABI tests)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fdn_b200.h")


def header_signatures():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(fdn_\w+)\s*\(([^)]*)\)\s*;", src):
        name, args = m.group(2), m.group(3).strip()
        codes = ""
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    codes += "p"
                elif a.startswith("cudaStream_t"):
                    codes += "s"
                elif a.startswith("long long"):
                    codes += "l"
                elif a.startswith("float"):
                    codes += "f"
                elif a.startswith("int"):
                    codes += "i"
                else:
                    raise ValueError("unparsed argument %r in %s" % (a, name))
        out[name] = codes
    return out
