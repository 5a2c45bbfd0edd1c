"""Regenerates report_tables.json from the reference's report.pdf (runs only where /root/reference exists).

The PDF has no text layer tools available in this image, so the content streams are inflated with
zlib and the TJ/Tj string operands concatenated; the two convergence tables are then read off the
token stream.  The committed JSON is the output of this script.
"""
import json
import re
import sys
import zlib

PDF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/report.pdf"


def page_texts(path):
    data = open(path, "rb").read()
    out = []
    for m in re.finditer(rb"stream\r?\n", data):
        s = m.end()
        e = data.find(b"endstream", s)
        try:
            d = zlib.decompress(data[s:e])
        except Exception:
            continue
        if b"TJ" not in d and b"Tj" not in d:
            continue
        txt = []
        for t in re.finditer(rb"\[(.*?)\]\s*TJ|\((.*?)\)\s*Tj", d, re.S):
            if t.group(1) is not None:
                txt.append(b"".join(re.findall(rb"\((.*?)(?<!\\)\)", t.group(1), re.S)))
            else:
                txt.append(t.group(2))
        out.append(b" ".join(txt).decode("latin1"))
    return out


def rows(text, start, stop):
    seg = text[text.index(start):]
    seg = seg[:seg.index(stop)] if stop in seg else seg
    toks = seg.split()
    out = []
    i = 0
    num = re.compile(r"^\d\.\d{3}e[+-]\d\d$")
    while i < len(toks):
        # row: ref 1 <hden> 1 <dtden> gmres rate err rate err rate err rate err rate
        if num.match(toks[i]):
            # walk back to the gmres count (token before optional rate)
            j = i - 1
            back = [toks[j], toks[j - 1]]
            gm = int(back[1]) if re.match(r"^-?\d+\.\d+$|^-$", back[0]) else int(back[0])
            errs = [toks[i]]
            k = i + 1
            while len(errs) < 4 and k < len(toks):
                if num.match(toks[k]):
                    errs.append(toks[k])
                k += 1
            out.append({"gmres": gm, "velocity_l2l2": float(errs[0]), "pressure_dg": float(errs[1]),
                        "pressure_l2l2": float(errs[2]), "lambda_l2": float(errs[3])})
            i = k
        else:
            i += 1
    return out


if __name__ == "__main__":
    txt = " ".join(page_texts(PDF))
    t2 = rows(txt, "Table2:", "2.2Quadratic")
    t4 = rows(txt, "Table4:", "3Plots")
    doc = {
        "source": "report.pdf (reference repository), Tables 2 and 4; 2x2 subdomains, patterns 3:2:4:3, mortar 1, "
                  "T=0.5, tol 1e-6, max 500 iterations",
        "table2_linear_mortar": {"cycles": [0, 1, 2, 3, 4], "rows": t2},
        "table4_quadratic_mortar": {"cycles": [0, 2, 4], "rows": t4},
    }
    json.dump(doc, sys.stdout, indent=1)
    print()
