"""Harvest the reference's known-answer vectors for the index decoders into index_kat.json.

Run in the build container (needs /root/reference):  python tests/golden/harvest_index_kat.py
Sources (reference demo/tests.cpp): :24-43 kIndexBuffer / kIndexDataV0 / kIndexBufferTricky / kIndexDataV1 /
kIndexSequence / kIndexSequenceV1 (used by decodeIndexV0 :80, decodeIndexV1 :91, decodeIndexSequence :277),
:102-134 decodeIndexV1More / decodeIndexV1ThreeEdges (input + ib), :243-256 decodeIndexMalformedVByte.
Only literal test DATA is extracted (array initialisers); no reference code is copied.
"""
import json, os, re, struct

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def c_array(text, name):
    m = re.search(r"\b" + re.escape(name) + r"\s*\[[^\]]*\]\s*=\s*\{(.*?)\};", text, re.S)
    assert m, name
    body = re.sub(r"//.*", "", m.group(1))
    return [int(t, 0) for t in re.findall(r"0x[0-9a-fA-F]+|\d+", body)]


def c_func(text, name):
    m = re.search(r"static void " + name + r"\(\)\n\{(.*?)\n\}\n", text, re.S)
    assert m, name
    return m.group(1)


def main():
    cpp = open(os.path.join(REF, "demo/tests.cpp")).read()
    kat = []

    def add(name, kind, data, expected, rc=0, count=None):
        kat.append({"name": "tests.cpp:" + name, "kind": kind, "count": len(expected) if count is None else count,
                    "input": bytes(data).hex(), "expected": struct.pack("<%dI" % len(expected), *expected).hex(), "rc": rc})

    add("decodeIndexV0", "triangles", c_array(cpp, "kIndexDataV0"), c_array(cpp, "kIndexBuffer"))
    add("decodeIndexV1", "triangles", c_array(cpp, "kIndexDataV1"), c_array(cpp, "kIndexBufferTricky"))
    for fn in ("decodeIndexV1More", "decodeIndexV1ThreeEdges"):
        body = c_func(cpp, fn)
        add(fn, "triangles", c_array(body, "input"), c_array(body, "ib"))
    add("decodeIndexSequence", "sequence", c_array(cpp, "kIndexSequenceV1"), c_array(cpp, "kIndexSequence"))
    add("decodeIndexMalformedVByte", "triangles", c_array(c_func(cpp, "decodeIndexMalformedVByte"), "input"), [], rc="negative", count=66)

    with open(os.path.join(HERE, "index_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote", len(kat), "vectors")


if __name__ == "__main__":
    main()
