"""Harvest the reference's own known-answer vectors for the vertex decode path into kat.json.

Run in the build container (needs /root/reference):  python tests/golden/harvest_kat.py
Sources (reference file:line):
  demo/tests.cpp:45-78     kVertexBuffer / kVertexDataV0 / kVertexDataV1 / kVertexDataV1Custom
  demo/tests.cpp:414-499   decodeVertexV0More / V0Mode2 / V1Deltas (input + expected)
  demo/tests.cpp:626-649   decodeVertexBitGroupSentinelCount
  demo/tests.cpp:762-876   decodeFilterOct8 / Oct12 / Quat12 / Exp (data + expected)
  js/meshopt_decoder.test.js:10-136   codec vectors incl. decodeVertexBufferV1_BitXorRotate
  js/meshopt_decoder.test.js:217-308  fused decode+filter vectors incl. Color8 / Color12
Only literal test DATA is extracted (array initialisers); no reference code is copied.
"""
import json, os, re, struct, sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def c_array(text, name, elem="B"):
    m = re.search(r"\b" + re.escape(name) + r"\s*\[[^\]]*\]\s*=\s*\{(.*?)\};", text, re.S)
    assert m, name
    body = re.sub(r"//.*", "", m.group(1))
    vals = [int(t, 0) for t in re.findall(r"0x[0-9a-fA-F]+|\d+", body)]
    return vals


def c_func(text, name):
    m = re.search(r"static void " + name + r"\(\)\n\{(.*?)\n\}\n", text, re.S)
    assert m, name
    return m.group(1)


def js_test(text, name):
    m = re.search(r"\n\t" + name + r": function \(\) \{(.*?)\n\t\},", text, re.S)
    assert m, name
    return m.group(1)


def js_array(body, var):
    m = re.search(r"var " + var + r" = new (Uint8Array|Uint16Array|Uint32Array)\(\[(.*?)\]\)", body, re.S)
    assert m, var
    kind = m.group(1)
    vals = [int(t, 0) for t in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(2))]
    fmt = {"Uint8Array": "B", "Uint16Array": "H", "Uint32Array": "I"}[kind]
    return struct.pack("<%d%s" % (len(vals), fmt), *vals)


def main():
    cpp = open(os.path.join(REF, "demo/tests.cpp")).read()
    js = open(os.path.join(REF, "js/meshopt_decoder.test.js")).read()
    kat = {"codec": [], "filter": [], "fused": []}

    # --- C++ codec vectors -------------------------------------------------------------------
    pv = re.search(r"static const PV kVertexBuffer\[\] = \{(.*?)\};", cpp, re.S).group(1)
    rows = re.findall(r"\{([^}]*)\}", pv)
    vb = b""
    for r in rows:
        px, py, pz, nu, nv, tx, ty = [int(t) for t in r.split(",")]
        vb += struct.pack("<HHHBBHH", px, py, pz, nu, nv, tx, ty)
    for name in ("kVertexDataV0", "kVertexDataV1", "kVertexDataV1Custom"):
        data = bytes(c_array(cpp, name))
        kat["codec"].append({"name": "tests.cpp:" + name, "count": 4, "size": 12, "input": data.hex(), "expected": vb.hex(), "rc": 0})

    for fn, count, size, efmt in (("decodeVertexV0More", 16, 4, "B"), ("decodeVertexV0Mode2", 16, 4, "B"), ("decodeVertexV1Deltas", 16, 8, "H"),
                                  ("decodeVertexBitGroupSentinelCount", 13, 4, "B")):
        body = c_func(cpp, fn)
        exp = c_array(body, "expected")
        inp = bytes(c_array(body, "input"))
        expb = struct.pack("<%d%s" % (len(exp), efmt), *exp)
        assert len(expb) == count * size, (fn, len(expb))
        kat["codec"].append({"name": "tests.cpp:" + fn, "count": count, "size": size, "input": inp.hex(), "expected": expb.hex(), "rc": 0})

    # --- C++ filter vectors ------------------------------------------------------------------------
    for fn, filt, stride, fmt in (("decodeFilterOct8", "oct", 4, "B"), ("decodeFilterOct12", "oct", 8, "H"),
                                  ("decodeFilterQuat12", "quat", 8, "H"), ("decodeFilterExp", "exp", 4, "I")):
        body = c_func(cpp, fn)
        data = c_array(body, "data")
        exp = c_array(body, "expected")
        kat["filter"].append({"name": "tests.cpp:" + fn, "filter": filt, "stride": stride, "count": 4,
                              "input": struct.pack("<%d%s" % (len(data), fmt), *data).hex(),
                              "expected": struct.pack("<%d%s" % (len(exp), fmt), *exp).hex()})

    # --- JS codec vectors --------------------------------------------------------------------------
    for fn, count, size in (("decodeVertexBuffer", 4, 12), ("decodeVertexBuffer_More", 16, 4), ("decodeVertexBuffer_Mode2", 16, 4),
                            ("decodeVertexBufferV1", 4, 12), ("decodeVertexBufferV1_Custom", 4, 12), ("decodeVertexBufferV1_Deltas", 16, 8),
                            ("decodeVertexBufferV1_BitXorRotate", 4, 16)):
        body = js_test(js, fn)
        inp, exp = js_array(body, "encoded"), js_array(body, "expected")
        assert len(exp) == count * size, fn
        kat["codec"].append({"name": "test.js:" + fn, "count": count, "size": size, "input": inp.hex(), "expected": exp.hex(), "rc": 0})

    # --- JS fused decode+filter vectors ------------------------------------------------------------
    for fn, count, size, filt in (("decodeFilterOct8", 4, 4, "oct"), ("decodeFilterOct12", 4, 8, "oct"), ("decodeFilterQuat12", 4, 8, "quat"),
                                  ("decodeFilterExp", 1, 16, "exp"), ("decodeFilterColor8", 4, 4, "color"), ("decodeFilterColor12", 4, 8, "color")):
        body = js_test(js, fn)
        inp, exp = js_array(body, "encoded"), js_array(body, "expected")
        assert len(exp) == count * size, fn
        kat["fused"].append({"name": "test.js:" + fn, "count": count, "size": size, "filter": filt, "input": inp.hex(), "expected": exp.hex()})

    # --- version probe vectors (demo/tests.cpp:743-751) ---------------------------------------------
    kat["version"] = [{"input": "a0", "rc": 0}, {"input": "a1", "rc": 1}, {"input": "a168656c6c6f", "rc": 1},
                      {"input": "", "rc": -1}, {"input": "a7", "rc": -1}, {"input": "b1", "rc": -1}]

    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print({k: len(v) for k, v in kat.items()})


if __name__ == "__main__":
    sys.exit(main())
