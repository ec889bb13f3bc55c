"""CPU restatement of the pipeline glue between the models (demo_image.py / find_embedding.py / data_loader).

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import os

import numpy as np
import torch

from . import align, detect, nets


def transforms_default(face_u8_hwc):
    """data_loader/__init__.py:27-34, 52-56: np.float32 -> (x-127.5)/128 -> HWC->CHW tensor."""
    arr = (np.float32(face_u8_hwc) - 127.5) / 128
    return torch.from_numpy(np.ascontiguousarray(np.transpose(arr, (2, 0, 1))))


def parallel_detect_and_align(rgb_images, sds, center_point, target_fs, min_face_size=50, faithful=True):
    """demo_image.py:273-306 with MTCNN(**cfg/detection/mtcnn.json) (image_size 160, keep_all, min_face_size 50,
    select_largest default True).  Returns (list[list[u8 (S,S,3)]], list[list[box]])."""
    import cv2
    boxes_b, _, lms_b = detect.mtcnn_detect(np.stack(rgb_images), sds, min_face_size=min_face_size, faithful=faithful)
    out_faces, out_boxes = [], []
    for img, boxes, lms in zip(rgb_images, boxes_b, lms_b):
        faces_i, chosen = [], []
        if len(boxes) > 0:
            crops, idx = align.get_face_from_boxes(img, boxes)
            chosen = [boxes[k] for k in idx]
            for c, k in zip(crops, idx):
                moved = lms[k] - boxes[k][:2]                       # move_landmark_to_box, demo_image.py:236-239
                bgr = cv2.cvtColor(c, cv2.COLOR_RGB2BGR)
                al = align.alignment(bgr, center_point, moved, target_fs[0], target_fs[1])
                faces_i.append(cv2.cvtColor(al, cv2.COLOR_BGR2RGB))
        out_faces.append(faces_i)
        out_boxes.append(chosen)
    return out_faces, out_boxes


def identify(emb, mlp_sd, threshold):
    """identify_person, demo_image.py:113-135, up to the label decision: argmax, exp(log-prob), threshold ->
    label or num_classes ("Unknown").  Returns (labels int64, probs fp32)."""
    with torch.no_grad():
        out = nets.mlp_forward(mlp_sd, emb)
    pred = torch.argmax(out, dim=1).numpy()
    prob = torch.exp(out).numpy()[np.arange(len(pred)), pred]
    n_classes = out.shape[1]
    lab = np.where(prob >= threshold, pred, n_classes)
    return lab.astype(np.int64), prob.astype(np.float32), out.numpy()


def recognize(bth_faces, enc_sd, mlp_sd, threshold=0.0, return_logp=False):
    """recognize_celeb, demo_image.py:50-76, without the name lookup: returns per-frame label lists + embeddings
    (+ the (F, C) log-probabilities when ``return_logp``)."""
    flat = [f for x in bth_faces for f in x]
    if not flat:
        empty = [[] for _ in bth_faces], np.zeros((0, 512), np.float32)
        return empty + (np.zeros((0, 0), np.float32),) if return_logp else empty
    x = torch.stack([transforms_default(f) for f in flat], 0)
    with torch.no_grad():
        emb = nets.encoder_forward(enc_sd, x)
    lab, prob, logp = identify(emb, mlp_sd, threshold)
    out, c = [], 0
    for faces in bth_faces:
        out.append(lab[c:c + len(faces)].tolist())
        c += len(faces)
    if return_logp:
        return out, emb.numpy(), logp
    return out, emb.numpy()


def cal_embedding(data_dir, batch_size, enc_sd, transforms, output_dir):
    """find_embedding.py:45-59 including the trailing-batch quirk of create_batch_images (:11-20): floor(n/bz) full
    batches plus ONE trailing batch, which the reference feeds to torch.stack even when empty (raises)."""
    from PIL import Image
    os.makedirs(output_dir, exist_ok=True)
    files = sorted(os.listdir(data_dir))
    nb = len(files) // batch_size
    batches = [files[i * batch_size:(i + 1) * batch_size] for i in range(nb)] + [files[nb * batch_size:]]
    for bf in batches:
        x = torch.stack([transforms(Image.open(os.path.join(data_dir, f))) for f in bf], 0)
        with torch.no_grad():
            emb = nets.encoder_forward(enc_sd, x).numpy()
        for i, f in enumerate(bf):
            np.savez_compressed(os.path.join(output_dir, f.split(".")[0] + ".npz"), emb[i])
