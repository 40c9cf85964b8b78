"""Round-2 golden fixtures, generated in the authoring container (needs /root/reference mounted):

    python tests/golden/make_golden_r2.py

  test_images.npz        YoloSharpDemo/Assets/TestImage/*.jpg decoded to uint8 CHW RGB (torchvision.io.read_image)
  yolov8n-seg_f16.npz    the reference's shipped yolov8n-seg.bin / yolov11n.bin re-encoded as npz (data fixtures)
  yolov11n_f16.npz
  v8n_images.npz         per test image: post-NMS rows + kept anchors of the fp32 oracle AND of the fp16-emulating
                         oracle (oracle/emul16.py) through Detector.ImagePredict semantics (conf 0.3, iou 0.7)
  v8nseg_bus.npz         Segmenter.ImagePredict on bus.jpg with yolov8n-seg.bin: rows, per-mask pixel counts and a
                         packed copy of the masks (fp32 oracle)

The reference itself is C#/TorchSharp and cannot be executed here, so these pin the ORACLE on the reference's
shipped assets; the only external anchor remains the README screenshot result for bus.jpg.
"""
import glob
import os
import sys

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import binfmt, emul16, ops, yolo  # noqa: E402

REF = "/root/reference/YoloSharpDemo/Assets/"


def main():
    imgs = {}
    for f in sorted(glob.glob(REF + "TestImage/*.jpg")):
        name = os.path.splitext(os.path.basename(f))[0]
        imgs[name] = torchvision.io.read_image(f, torchvision.io.ImageReadMode.RGB)
    np.savez_compressed(os.path.join(HERE, "test_images.npz"), **{k: v.numpy() for k, v in imgs.items()})
    for bin_name, out_name in (("yolov8n-seg.bin", "yolov8n-seg_f16.npz"), ("yolov11n.bin", "yolov11n_f16.npz")):
        sd, trailing = binfmt.load_bin(REF + "PreTrainedModels/" + bin_name)
        assert trailing == 0
        np.savez_compressed(os.path.join(HERE, out_name), **{k: v.numpy() for k, v in sd.items()})

    m = yolo.build("v8", "detect", "n").eval()
    binfmt.load_into(m, REF + "PreTrainedModels/Yolov8n.bin")
    me = emul16.convert(m)
    z = {}
    for name, img in imgs.items():
        rows, keep, _ = ops.detector_predict(m, img, 0.3, 0.7)
        z[name + "_rows"], z[name + "_keep"] = rows.numpy(), keep.numpy()
        pre16 = lambda im: emul16.input_u8(ops.preprocess(im) * 255.0)  # noqa: E731  (pad-114 then the stem's u8 scaling)
        rows, keep, _ = ops.detector_predict(me, img, 0.3, 0.7, pre=pre16)
        z[name + "_rows16"], z[name + "_keep16"] = rows.numpy(), keep.numpy()
        print(name, tuple(img.shape), "fp32:", z[name + "_rows"].shape[0], "fp16-emul:", rows.shape[0])
    np.savez_compressed(os.path.join(HERE, "v8n_images.npz"), **z)

    ms = yolo.build("v8", "segment", "n").eval()
    binfmt.load_into(ms, REF + "PreTrainedModels/yolov8n-seg.bin")
    rows, masks, res = ops.segmenter_predict(ms, imgs["bus"], 0.3, 0.7)
    np.savez_compressed(os.path.join(HERE, "v8nseg_bus.npz"), rows=rows.numpy(), mask_pixels=masks.sum((1, 2)).numpy(),
                        masks_packed=np.packbits(masks.numpy().astype(bool), axis=-1), mask_shape=np.array(masks.shape))
    print("seg bus:", [(r["ClassID"], round(r["Score"], 3)) for r in res], masks.sum((1, 2)).tolist())


if __name__ == "__main__":
    main()
