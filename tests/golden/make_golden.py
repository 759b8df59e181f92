"""Generates the golden fixtures under tests/golden/ from the reference's SHIPPED DATA (run in the
build container, where /root/reference is mounted; the GPU box never sees /root/reference).

    python tests/golden/make_golden.py

Decodes gmmloc_ros/data/map/v{1,2}.gmm with an independent pure-Python reader of the wire format
(varint32 count, then per component varint32 size + ComponentProto bytes: gmmloc/src/utils/
protobuf_utils.cpp:12-80, gmmloc/proto/gmmloc/GMM.proto:5-14) and stores

  v1_gmm.npz / v2_gmm.npz : mean Kx3, cov Kx9 (row-major as the loader fills it, gmm_utils.cpp:55-59),
                            flags Kx2 = the STORED is_degenerated / is_salient bools (the one implicit
                            golden vector of the reference: they were written by the reference's own
                            GaussianComponent::decompose, gaussian.cpp:39-55), sha256 + size of the file
  v1_03_gt_sync.npz       : the 2 149 camera poses of gmmloc_ros/data/gt_sync/V1_03_difficult.txt
                            (t x y z qx qy qz qw, T_wc; loader dataloader.cpp:118-138)
"""
import hashlib
import struct
from pathlib import Path

import numpy as np

REF = Path("/root/reference/gmmloc_ros/data")
OUT = Path(__file__).resolve().parent


def varint(b, pos):
    v, shift = 0, 0
    while True:
        c = b[pos]
        pos += 1
        v |= (c & 0x7F) << shift
        if not c & 0x80:
            return v, pos
        shift += 7


def decode_gmm(raw: bytes):
    pos = 0
    count, pos = varint(raw, pos)
    mean = np.zeros((count, 3))
    cov = np.zeros((count, 9))
    flags = np.zeros((count, 2), np.uint8)
    for k in range(count):
        size, pos = varint(raw, pos)
        end = pos + size
        while pos < end:
            tag, pos = varint(raw, pos)
            field, wt = tag >> 3, tag & 7
            if wt == 0:
                v, pos = varint(raw, pos)
                if field in (1, 2):
                    flags[k, field - 1] = 1 if v else 0
            elif wt == 2:
                ln, pos = varint(raw, pos)
                vals = struct.unpack("<%dd" % (ln // 8), raw[pos:pos + ln])
                if field == 3:
                    mean[k] = vals
                elif field == 4:
                    cov[k] = vals
                pos += ln
            else:
                raise ValueError("unexpected wire type %d" % wt)
        assert pos == end
    assert pos == len(raw)
    return mean, cov, flags


def main():
    for name in ("v1", "v2"):
        raw = (REF / "map" / f"{name}.gmm").read_bytes()
        mean, cov, flags = decode_gmm(raw)
        np.savez_compressed(OUT / f"{name}_gmm.npz", mean=mean, cov=cov, flags=flags,
                            sha256=np.array(hashlib.sha256(raw).hexdigest()), nbytes=np.array(len(raw)))
        print(name, mean.shape[0], "components,", int(flags[:, 0].sum()), "degenerate,", len(raw), "bytes")
    gt = np.loadtxt(REF / "gt_sync" / "V1_03_difficult.txt")
    np.savez_compressed(OUT / "v1_03_gt_sync.npz", gt=gt)
    print("V1_03 gt_sync", gt.shape)


if __name__ == "__main__":
    main()
