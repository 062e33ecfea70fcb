#!/usr/bin/env python
"""Golden vectors on the reference's OWN content (run once in the build container; needs /root/reference and oracle/_ref):

  * the three .agmv streams the reference ships (agmv_splash.agmv, agmv_spash.agmv - the one with the damaged header -
    and examples/simple_decoding/FOXLOGO.agmv) are copied to tests/golden/fixtures/ as decode-only test vectors; the
    unmodified reference decodes them here and the digests of its frames and audio track go to golden.json;
  * the first 20 frames of examples/simple_video/foxlogo/ are copied to tests/golden/foxlogo/ and encoded by the unmodified
    reference the way examples/simple_video/simple_video.c does (OPT_I, LOW, LZSS): the stream is the golden output of an
    encode on real (non-synthetic) BMP files.

    python tests/golden/make_fixture_golden.py
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from agmv_testlib import REF_DIR, ref_audio_track, ref_decode_raw, sha256  # noqa: E402

REF = "/root/reference"
N_FOX = 20


def main():
    gold_path = os.path.join(HERE, "golden.json")
    gold = json.load(open(gold_path))
    fx = gold.setdefault("fixtures", {})
    os.makedirs(os.path.join(HERE, "fixtures"), exist_ok=True)
    for rel in ("agmv_splash.agmv", "agmv_spash.agmv", "examples/simple_decoding/FOXLOGO.agmv"):
        name = os.path.basename(rel)
        data = open(os.path.join(REF, rel), "rb").read()
        shutil.copy(os.path.join(REF, rel), os.path.join(HERE, "fixtures", name))
        rc, frames = ref_decode_raw(data)
        e = dict(file=f"fixtures/{name}", input_sha256=sha256(data), size=len(data), rc=rc)
        if rc == 0:
            e["decoded_shape"] = list(frames.shape)
            e["decoded_sha256"] = sha256(frames.tobytes())
            e["decoded_frame_sha256"] = [sha256(f.tobytes()) for f in frames]
            arc, pcm = ref_audio_track(data)
            assert arc == 0
            e["audio"] = dict(bits=16 if pcm.dtype == np.uint16 else 8, samples=int(pcm.size), sha256=sha256(pcm.tobytes()))
        fx[name] = e
        print(name, rc, e.get("decoded_shape"), e.get("audio"))
    # real BMP input
    dst = os.path.join(HERE, "foxlogo")
    os.makedirs(dst, exist_ok=True)
    for k in range(1, N_FOX + 1):
        shutil.copy(os.path.join(REF, "examples/simple_video/foxlogo", f"foxlogo{k}.bmp"), os.path.join(dst, f"foxlogo{k}.bmp"))
    with tempfile.TemporaryDirectory() as td:
        # ref_encode OUT DIR BASE START END W H FPS OPT QUALITY COMPRESSION CREATE_N (paths are formatted into char[60]: keep them short)
        os.symlink(dst, os.path.join(td, "fx"))
        subprocess.run([os.path.join(REF_DIR, "ref_encode"), "o.agmv", "fx", "foxlogo", "1", str(N_FOX), "320", "240", "24", "1", "3", "1", str(N_FOX)],
                       cwd=td, check=True, stdout=subprocess.DEVNULL)
        data = open(os.path.join(td, "o.agmv"), "rb").read()
    open(os.path.join(HERE, "foxlogo_I_LOW.agmv"), "wb").write(data)
    rc, frames = ref_decode_raw(data)
    assert rc == 0
    fx["foxlogo_I_LOW"] = dict(file="foxlogo_I_LOW.agmv", dir="foxlogo", base="foxlogo", n=N_FOX, w=320, h=240, fps=24, opt="I", quality="LOW",
                               create_n=N_FOX, size=len(data), sha256=sha256(data), decoded_shape=list(frames.shape),
                               decoded_frame_sha256=[sha256(f.tobytes()) for f in frames])
    print("foxlogo", len(data), frames.shape)
    json.dump(gold, open(gold_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
