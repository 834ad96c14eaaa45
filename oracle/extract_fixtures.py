#!/usr/bin/env python3
"""Extract the reference's only data fixtures (samples/ba_input.7z) into flat binary graphs.

TEST INFRASTRUCTURE ONLY (oracle/): nothing in the product path imports this.

The archive is one LZMA2 folder (dict 64 MiB) starting at byte offset 32, packed size
12 500 131, holding ba_kitti_00.json (147 447 492 B, CRC32 0x4e080615) followed by
ba_kitti_07.json (25 560 898 B, CRC32 0x6367a643) -- SURVEY.md fact 4.  No 7z binary
is needed: python's stdlib lzma decodes the raw LZMA2 stream.

JSON schema (reference samples/sample_ba_from_file.cpp:93-157): pose_vertices[{id,fixed,q[xyzw],t}],
landmark_vertices[{id,fixed,Xw}], monocular_edges[{vertexP,vertexL,measurement[2],information}],
stereo_edges[{...measurement[3]...}], scalars fx fy cx cy bf.

Output: oracle/_ref/fixtures/<name>.cubagraph  (format: see cuda-bundle-adjustment_b200/graphio.py).
oracle/_ref/ is git-ignored (reference-derived data stays out of history) but travels with gpurun.
"""
import json
import lzma
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = "/root/reference/samples/ba_input.7z"
OUT = os.path.join(HERE, "_ref", "fixtures")

FILES = [("ba_kitti_00", 147447492, 0x4E080615), ("ba_kitti_07", 25560898, 0x6367A643)]
PACK_OFFSET, PACK_SIZE = 32, 12500131

MAGIC = b"CUBAGRF1"


def write_graph(path, g):
    """g: dict of numpy arrays, see graphio.py for the layout."""
    with open(path, "wb") as f:
        f.write(MAGIC)
        np.array([len(g["pose_id"]), len(g["lm_id"]), len(g["mono_vP"]), len(g["stereo_vP"])], dtype=np.int64).tofile(f)
        for key, dt in (("pose_id", np.int32), ("pose_fixed", np.int32), ("q", np.float64), ("t", np.float64),
                        ("cam", np.float64), ("lm_id", np.int32), ("lm_fixed", np.int32), ("Xw", np.float64),
                        ("mono_vP", np.int32), ("mono_vL", np.int32), ("mono_meas", np.float64), ("mono_info", np.float64),
                        ("stereo_vP", np.int32), ("stereo_vL", np.int32), ("stereo_meas", np.float64),
                        ("stereo_info", np.float64)):
            np.ascontiguousarray(g[key], dtype=dt).tofile(f)


def json_to_graph(doc):
    pv, lv = doc["pose_vertices"], doc["landmark_vertices"]
    me, se = doc["monocular_edges"], doc["stereo_edges"]
    cam = np.array([doc["fx"], doc["fy"], doc["cx"], doc["cy"], doc["bf"]], dtype=np.float64)
    g = {
        "pose_id": np.array([v["id"] for v in pv], dtype=np.int32),
        "pose_fixed": np.array([int(v["fixed"]) for v in pv], dtype=np.int32),
        "q": np.array([v["q"] for v in pv], dtype=np.float64).reshape(-1, 4),
        "t": np.array([v["t"] for v in pv], dtype=np.float64).reshape(-1, 3),
        "lm_id": np.array([v["id"] for v in lv], dtype=np.int32),
        "lm_fixed": np.array([int(v["fixed"]) for v in lv], dtype=np.int32),
        "Xw": np.array([v["Xw"] for v in lv], dtype=np.float64).reshape(-1, 3),
        "mono_vP": np.array([e["vertexP"] for e in me], dtype=np.int32),
        "mono_vL": np.array([e["vertexL"] for e in me], dtype=np.int32),
        "mono_meas": np.array([e["measurement"] for e in me], dtype=np.float64).reshape(-1, 2),
        "mono_info": np.array([e["information"] for e in me], dtype=np.float64),
        "stereo_vP": np.array([e["vertexP"] for e in se], dtype=np.int32),
        "stereo_vL": np.array([e["vertexL"] for e in se], dtype=np.int32),
        "stereo_meas": np.array([e["measurement"] for e in se], dtype=np.float64).reshape(-1, 3),
        "stereo_info": np.array([e["information"] for e in se], dtype=np.float64),
    }
    g["cam"] = np.tile(cam, (len(pv), 1))
    return g


def main():
    if not os.path.exists(ARCHIVE):
        print("extract_fixtures: %s absent (GPU box?) - nothing to do" % ARCHIVE)
        return 0
    os.makedirs(OUT, exist_ok=True)
    if all(os.path.exists(os.path.join(OUT, n + ".cubagraph")) for n, _, _ in FILES):
        print("extract_fixtures: fixtures already present")
        return 0
    raw = open(ARCHIVE, "rb").read()[PACK_OFFSET:PACK_OFFSET + PACK_SIZE]
    dec = lzma.LZMADecompressor(format=lzma.FORMAT_RAW,
                                filters=[{"id": lzma.FILTER_LZMA2, "dict_size": 1 << 26}])
    data = dec.decompress(raw)
    off = 0
    for name, size, crc in FILES:
        blob = data[off:off + size]
        off += size
        got = zlib.crc32(blob) & 0xFFFFFFFF
        if got != crc:
            raise SystemExit("CRC mismatch for %s: %08x != %08x" % (name, got, crc))
        g = json_to_graph(json.loads(blob))
        path = os.path.join(OUT, name + ".cubagraph")
        write_graph(path, g)
        print("wrote %s: %d poses %d landmarks %d mono %d stereo" %
              (path, len(g["pose_id"]), len(g["lm_id"]), len(g["mono_vP"]), len(g["stereo_vP"])))
    return 0


if __name__ == "__main__":
    sys.exit(main())
