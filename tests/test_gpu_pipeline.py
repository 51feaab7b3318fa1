"""DetectionPipeline (network + decode/NMS behind one call): host API and streaming API agree with the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from k210_yolo_framework_b200.pipeline import DetectionPipeline
from oracle import decode_ref, keras_ref


def _batch(dog_u8, n):
    x = (dog_u8 / np.max(dog_u8)).astype(np.float32)
    imgs = [x, x[::-1].copy(), x[:, ::-1].copy(), np.roll(x, 40, axis=1)]
    return np.stack([imgs[i % 4] for i in range(n)])


def test_host_and_streaming_api_match_oracle(golden_weights, voc_anchors, dog_u8):
    n = 4
    pipe = DetectionPipeline("yolo_mobilev1", (224, 320), voc_anchors, 20, 0.75, n, obj_thresh=0.7, iou_thresh=0.5)
    pipe.engine.set_weights(golden_weights)
    x = _batch(dog_u8, n)
    xh = torch.from_numpy(x).pin_memory()
    dets, counts = pipe.detect_host(xh)
    got = DetectionPipeline.records(dets.clone(), counts.clone())
    heads = keras_ref.forward("yolo_mobilev1", golden_weights, x, alpha=0.75)
    h = decode_ref.HelperRef(voc_anchors, [224, 320], [7, 10, 14, 20], 20)
    ref = decode_ref.detect_batch_fast(heads, h, [224, 320], [(224, 320)] * n, 0.7, 0.5)
    assert [(d[0], d[1]) for d in got[0]] == [(6, 53), (11, 765)]
    for g_img, r_img in zip(got, ref):
        assert [(d[0], d[1]) for d in g_img] == [(d[0], d[1]) for d in r_img]          # identical post-NMS indices
        np.testing.assert_allclose([d[2] for d in g_img], [float(d[2]) for d in r_img], atol=1e-3)
        np.testing.assert_allclose(np.array([d[3:] for d in g_img]).reshape(-1, 4) / [224, 320, 224, 320],
                                   np.array([[float(v) for v in d[3:]] for d in r_img]).reshape(-1, 4) / [224, 320, 224, 320],
                                   atol=1e-3)
    # streaming: three batches in flight two at a time; every ticket returns its own batch's records
    batches = [xh, torch.from_numpy(x[::-1].copy()).pin_memory(), xh]
    tickets, outs = [], []
    for i, bx in enumerate(batches):
        tickets.append(pipe.submit(bx))
        if i >= 1:
            d, c = pipe.collect(tickets[i - 1])
            outs.append(DetectionPipeline.records(d.clone(), c.clone()))
    d, c = pipe.collect(tickets[-1])
    outs.append(DetectionPipeline.records(d.clone(), c.clone()))
    assert outs[0] == got and outs[2] == got
    assert outs[1] == got[::-1]
    assert pipe.launches_per_step() >= 33


def test_uint8_front_end_is_bit_identical(golden_weights, voc_anchors, dog_u8):
    """uint8 input + on-GPU `img / np.max(img)` == feeding the float32 array the reference builds on the host."""
    n = 4
    pipe = DetectionPipeline("yolo_mobilev1", (224, 320), voc_anchors, 20, 0.75, n, obj_thresh=0.7, iou_thresh=0.5)
    pipe.engine.set_weights(golden_weights)
    imgs = [dog_u8, dog_u8[::-1].copy(), (dog_u8 // 2).astype(np.uint8), np.roll(dog_u8, 40, axis=1)]   # image 2 has max 127
    u8 = np.stack(imgs)
    f32 = np.stack([(im / np.max(im)).astype(np.float32) for im in imgs])
    heads_f = [t.clone() for t in pipe.engine.predict_device(torch.from_numpy(f32).cuda())]
    heads_u = [t.clone() for t in pipe.engine.predict_device_u8(torch.from_numpy(u8).cuda())]
    for a, b in zip(heads_f, heads_u):
        assert torch.equal(a, b)
    d_u, c_u = pipe.detect_host(torch.from_numpy(u8).pin_memory())
    got_u = DetectionPipeline.records(d_u.clone(), c_u.clone())
    d_f, c_f = pipe.detect_host(torch.from_numpy(f32).pin_memory())
    assert DetectionPipeline.records(d_f.clone(), c_f.clone()) == got_u
    assert [(d[0], d[1]) for d in got_u[0]] == [(6, 53), (11, 765)]
    tk = pipe.submit(torch.from_numpy(u8).pin_memory())
    d_s, c_s = pipe.collect(tk)
    assert DetectionPipeline.records(d_s.clone(), c_s.clone()) == got_u
