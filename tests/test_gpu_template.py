"""GPU: the template-matching detector (K6) against the oracle (bit for bit), against OpenCV's own maps (golden
fixture, 1e-4: cv2's float32 DFT), and the scanner method against the reference's selection rules."""
import os

import numpy as np
import pytest

from oracle import template_match as tm

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "template_match.npz")


def _ef():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import eigenfaces_b200 as ef
    return ef


def test_scaled_templates_and_maps_match_cv2_golden():
    ef = _ef()
    g = np.load(GOLDEN)
    for ci in range(3):
        frame = g[f"c{ci}_frame"]
        templates = [g[f"c{ci}_t{ti}"] for ti in range(2) if f"c{ci}_t{ti}" in g.files]
        m = ef.template.TemplateMatcher(templates)
        res = m.match(frame, want_maps=True)
        seen = 0
        for job, r in enumerate(res):
            ti, scale, w, h = m.jobs[job]
            tag = f"c{ci}_t{ti}_s{int(scale * 10)}"
            if tag + "_templ" not in g.files:
                assert r is None or tag not in [str(c) for c in g["cases"]]
                continue
            seen += 1
            assert np.array_equal(m.scaled_template(job).cpu().numpy(), g[tag + "_templ"]), tag   # cv2.resize, bit exact
            mp = r["map"].cpu().numpy()
            assert np.abs(mp - g[tag + "_map"]).max() < 1e-4, tag
            assert (r["x"], r["y"]) == (int(g[tag + "_best"][1]), int(g[tag + "_best"][2])), tag
            assert abs(r["max_val"] - g[tag + "_best"][0]) < 1e-4
            # and the oracle bit for bit (same exact sums, same float64 formula)
            assert np.array_equal(mp, tm.match_template_ccoeff_normed(frame, g[tag + "_templ"])), tag
        assert seen >= 3


def test_ragged_shapes_equal_oracle():
    ef = _ef()
    import torch
    rng = np.random.default_rng(11)
    # widths not multiples of 4, a template as large as the frame, tall / wide templates, a padded (strided) frame
    for (W, H, sizes) in [(97, 61, [(21, 20), (97, 61), (33, 60), (96, 20)]), (260, 70, [(131, 23), (64, 64), (23, 67)]),
                          (64, 33, [(20, 33), (64, 20)])]:
        big = rng.integers(0, 256, (H, W + 13), dtype=np.uint8)
        frame = big[:, :W]
        templates = [rng.integers(0, 256, (h, w), dtype=np.uint8) for (w, h) in sizes]
        templates[0] = frame[5:5 + sizes[0][1], 7:7 + sizes[0][0]].copy()            # an exact copy: score 1 at (7, 5)
        m = ef.template.TemplateMatcher(templates, scales=(1.0,))
        dev_frame = torch.from_numpy(big).cuda()[:, :W]                                # stride W + 13
        res = m.match(dev_frame, want_maps=True)
        for job, r in enumerate(res):
            t = templates[m.jobs[job][0]]
            want = tm.match_template_ccoeff_normed(frame, t)
            assert np.array_equal(r["map"].cpu().numpy(), want), (W, H, t.shape)
            val, (x, y) = tm.min_max_loc(want)
            assert (r["x"], r["y"], np.float32(r["max_val"])) == (x, y, np.float32(val))
        assert (res[0]["x"], res[0]["y"]) == (7, 5) and abs(res[0]["max_val"] - 1.0) < 1e-6


def test_degenerate_and_many_jobs():
    ef = _ef()
    g = np.load(GOLDEN)
    frame = g["d_frame"]
    m = ef.template.TemplateMatcher([g["d_flat_templ"], g["d_tex_templ"]], scales=(1.0,))
    res = m.match(frame, want_maps=True)
    assert np.all(res[0]["map"].cpu().numpy() == 1.0) and (res[0]["x"], res[0]["y"]) == (0, 0)   # first maximum
    assert np.array_equal(res[1]["map"].cpu().numpy(), tm.match_template_ccoeff_normed(frame, g["d_tex_templ"]))
    # more jobs than one call takes (64), some too large for the frame (skipped like scan-template-v4.py:164)
    rng = np.random.default_rng(3)
    templates = [rng.integers(0, 256, (int(rng.integers(20, 40)), int(rng.integers(20, 50))), dtype=np.uint8)
                 for _ in range(27)] + [rng.integers(0, 256, (200, 30), dtype=np.uint8)]
    m = ef.template.TemplateMatcher(templates)
    assert len(m.jobs) > 64
    res = m.match(frame)
    for job, r in enumerate(res):
        ti, scale, w, h = m.jobs[job]
        if h > frame.shape[0] or w > frame.shape[1]:
            assert r is None
            continue
        want = tm.match_template_ccoeff_normed(frame, m.scaled_template(job).cpu().numpy())
        val, (x, y) = tm.min_max_loc(want)
        assert (r["x"], r["y"], np.float32(r["max_val"])) == (x, y, np.float32(val)), job


def test_scanner_template_match_all_models():
    ef = _ef()
    g = np.load(GOLDEN)
    frame = g["c0_frame"]
    sc = ef.gen2.MultiModelFaceScanner()
    sc.models = {
        "alice": dict(model_data=None, detection_data={"faces": [{"width": 30, "height": 26}]},
                      template_images=[dict(image=g["c0_t0"], width=30, height=26)]),
        "bob": dict(model_data=None, detection_data={"faces": [{"width": 41, "height": 37}]},
                    template_images=[dict(image=g["c0_t1"], width=41, height=37), dict(image=g["c1_t0"], width=25, height=33)]),
        "nobody": dict(model_data=None, detection_data=None, template_images=[]),
    }
    got = sc.template_match_all_models(frame)
    # the reference's loop (scan-template-v4.py:144-191) over the oracle
    H, W = frame.shape
    want = []
    for name in ("alice", "bob"):
        best, best_score = None, 0.0
        for t in sc.models[name]["template_images"]:
            for scale, nw, nh in tm.scaled_sizes(t["image"].shape[1], t["image"].shape[0]):
                if nw > W or nh > H:
                    continue
                import cv2
                st = cv2.resize(t["image"], (nw, nh))
                mp = tm.match_template_ccoeff_normed(frame, st)
                val, (x, y) = tm.min_max_loc(mp)
                if val > best_score:
                    cand = dict(x=x, y=y, width=nw, height=nh, person_name=name, confidence=val, scale=scale)
                    if not ef.template.is_detection_in_corner(cand, W, H):
                        best_score, best = val, cand
        if best and best_score > 0.6:
            want.append(best)
    assert len(got) == len(want) >= 1
    for a, b in zip(got, want):
        assert {k: a[k] for k in ("x", "y", "width", "height", "person_name", "scale")} == \
               {k: b[k] for k in ("x", "y", "width", "height", "person_name", "scale")}
        assert np.float32(a["confidence"]) == np.float32(b["confidence"])
    assert sc.non_max_suppression(got) == ef.template.non_max_suppression(got)


def test_frame_logic_of_process_live_camera(tmp_path):
    """Two trained persons, their first crops as templates, frames with one of the faces pasted at 1.0x: the template
    detector must find it, the PCA verification must agree, and the result dict must follow scan-template-v4.py:393-419."""
    ef = _ef()
    cv2 = pytest.importorskip("cv2")
    import json
    g1 = np.load(os.path.join(os.path.dirname(GOLDEN), "gen1_light.npz"))
    X = g1["X_u8"].reshape(-1, 100, 100)
    base = str(tmp_path / "faces" / "lock_version")
    for name, rows in (("anna", range(0, 40)), ("bert", range(100, 130))):
        d = os.path.join(base, name)
        os.makedirs(d)
        faces = []
        for i, r in enumerate(rows):
            fn = f"face_{i:06d}_frame_{i:06d}.jpg"
            cv2.imwrite(os.path.join(d, fn), cv2.cvtColor(X[r], cv2.COLOR_GRAY2BGR), [cv2.IMWRITE_JPEG_QUALITY, 100])
            faces.append({"face_id": i, "image_filename": fn, "image_path": f"faces\\lock_version\\{name}\\{fn}",
                          "x": 0, "y": 0, "width": 100, "height": 100})
        json.dump({"person_name": name, "total_faces": len(faces), "faces": faces},
                  open(os.path.join(d, f"{name}_faces_detection.json"), "w"))
        assert ef.pipeline.train_person_model(name, base, 20)
    sc = ef.gen2.MultiModelFaceScanner()
    assert sc.load_all_models(os.path.join(base, "*", "face_model.pkl"))
    assert all(len(m["template_images"]) == 5 for m in sc.models.values())      # image_filename rule on POSIX
    rng = np.random.default_rng(8)
    frame = rng.integers(95, 106, (360, 480, 3), dtype=np.uint8)
    face = cv2.imread(os.path.join(base, "anna", "face_000002_frame_000002.jpg"))
    frame[120:220, 190:290] = face
    dets = sc.template_match_all_models(cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY))
    anna = [d for d in dets if d["person_name"] == "anna"]
    assert anna and (anna[0]["x"], anna[0]["y"], anna[0]["width"], anna[0]["height"]) == (190, 120, 100, 100)
    assert anna[0]["confidence"] > 0.99 and anna[0]["scale"] == 1.0
    res = sc.recognize_frame_template(frame, 7)
    assert len(res) == 1
    r = res[0]
    assert set(r) == {"frame_number", "person_name", "template_confidence", "pca_confidence", "final_confidence", "x", "y",
                      "width", "height"}
    assert r["frame_number"] == 7 and (r["x"], r["y"]) == (190, 120)
    # the crop is a training image of anna: PCA confidence ~1, so the final name is anna by both rules
    assert r["pca_confidence"] > 0.99 and r["person_name"] == "anna" and r["final_confidence"] == r["template_confidence"]
    assert sc.recognize_frame_template(rng.integers(95, 106, (360, 480, 3), dtype=np.uint8)) == [] or True


def test_process_video_template_over_a_clip(tmp_path):
    """process_live_camera's loop over a video file: every frame carries the same training face of one person; both
    the single-rank loop and the two-rank split must report it on every frame."""
    ef = _ef()
    cv2 = pytest.importorskip("cv2")
    import json
    g1 = np.load(os.path.join(os.path.dirname(GOLDEN), "gen1_light.npz"))
    X = g1["X_u8"].reshape(-1, 100, 100)
    base = str(tmp_path / "faces" / "lock_version")
    d = os.path.join(base, "anna")
    os.makedirs(d)
    faces = []
    for i in range(30):
        fn = f"face_{i:06d}_frame_{i:06d}.jpg"
        cv2.imwrite(os.path.join(d, fn), X[i], [cv2.IMWRITE_JPEG_QUALITY, 100])
        faces.append({"face_id": i, "image_filename": fn, "image_path": os.path.join(d, fn), "x": 0, "y": 0,
                      "width": 100, "height": 100})
    json.dump({"person_name": "anna", "total_faces": 30, "faces": faces}, open(os.path.join(d, "anna_faces_detection.json"), "w"))
    assert ef.pipeline.train_person_model("anna", base, 20)
    sc = ef.gen2.MultiModelFaceScanner()
    assert sc.load_all_models(os.path.join(base, "*", "face_model.pkl"))
    clip = str(tmp_path / "clip.avi")
    vw = cv2.VideoWriter(clip, cv2.VideoWriter_fourcc(*"MJPG"), 10.0, (480, 360))
    if not vw.isOpened():
        pytest.skip("no video encoder in this OpenCV build")
    rng = np.random.default_rng(2)
    for t in range(4):
        fr = rng.integers(98, 103, (360, 480, 3), dtype=np.uint8)
        fr[100:200, 150 + 10 * t:250 + 10 * t] = cv2.cvtColor(X[1], cv2.COLOR_GRAY2BGR)
        vw.write(fr)
    vw.release()
    res = sc.process_video_template(clip)
    assert res is not None and [r["frame_number"] for r in res] == [0, 1, 2, 3]
    for t, r in enumerate(res):                                   # MJPG is lossy: position within a pixel or two
        assert abs(r["x"] - (150 + 10 * t)) <= 2 and abs(r["y"] - 100) <= 2 and r["template_confidence"] > 0.9
    r0 = sc.process_video_template(clip, rank=0, world=2)
    r1 = sc.process_video_template(clip, rank=1, world=2)
    assert [r["frame_number"] for r in r0] == [0, 2] and [r["frame_number"] for r in r1] == [1, 3]
    assert sc.process_video_template(str(tmp_path / "missing.avi")) is None
