"""Generates the golden fixtures in this directory.  Run ONCE in the build container (where
/root/reference is mounted); the tests only read the committed outputs.

What can be pinned against the real reference here (no TensorFlow / librosa / soundfile in the image):
  * evaluation/pooling.py           -> pooling_reference.npz   (imported as is)
  * audio/io.py chunk geometry      -> chunking_reference.json (imported with a stub `soundfile` module:
                                       split_audio_into_chunks / estimate_num_chunks never touch it)
  * training/config.py              -> config_reference.json   (legacy JSON defaulting)
Self-generated (oracle outputs, to detect drift of the oracle itself; NOT reference outputs):
  * oracle_selfcheck.npz            scores / taps of the CPU oracle on the seeded 8-chunk batch
"""

import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/birdnet_stm32"


def load_ref(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod          # dataclasses look the module up by name
    spec.loader.exec_module(mod)
    return mod


def main():
    # ---- pooling ---------------------------------------------------------------------------------
    pooling = load_ref("evaluation/pooling.py", "ref_pooling")
    rng = np.random.default_rng(20261017)
    out = {}
    cases = []
    for i, (n, c) in enumerate([(1, 5), (2, 2), (7, 100), (20, 100), (21, 37), (3, 1)]):
        s = rng.random((n, c)).astype(np.float32)
        if i == 3:
            s[:, :10] = np.round(s[:, :10] * 256) / 256      # LOGISTIC-like quantised scores
        out[f"in_{i}"] = s
        for method, beta in (("avg", 10.0), ("max", 10.0), ("lme", 10.0), ("lme", 1.0), ("lme", 50.0)):
            key = f"out_{i}_{method}_{beta:g}"
            out[key] = np.asarray(pooling.pool_scores(s, method=method, beta=beta))
            cases.append(key)
    out["empty_avg"] = pooling.pool_scores(np.zeros((0, 5), np.float32), method="avg")
    np.savez_compressed(os.path.join(HERE, "pooling_reference.npz"), **out)

    # ---- chunk geometry (stub soundfile) ---------------------------------------------------------------
    sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
    io = load_ref("audio/io.py", "ref_audio_io")
    geo = []
    for sr, cd in ((22050, 3.0), (24000, 3.0), (24000, 2.0), (16000, 3.0)):
        size = int(sr * cd)
        for n in (0, 1, size // 2, size - 1, size, size + 1, 2 * size, 2 * size + 17, int(7.5 * size), 20 * size, 20 * size + 5):
            for ov in (0.0, 0.5, 1.5, 2.95, 5.0):
                y = (np.arange(n) % 251).astype(np.float32)
                chunks = io.split_audio_into_chunks(y, sample_rate=sr, chunk_duration=cd, chunk_overlap=ov)
                firsts = [int(ch[0]) if ch.size else -1 for ch in chunks]
                lasts = [int(ch[-1]) if ch.size else -1 for ch in chunks]
                geo.append(dict(sr=sr, cd=cd, n=n, overlap=ov, n_chunks=int(chunks.shape[0]), shape1=int(chunks.shape[1]),
                                estimate=int(io.estimate_num_chunks(n, sr, cd, ov)), firsts=firsts, lasts=lasts))
    json.dump(geo, open(os.path.join(HERE, "chunking_reference.json"), "w"))

    # ---- config defaulting ------------------------------------------------------------------------------
    config = load_ref("training/config.py", "ref_config")
    shipped = config.ModelConfig.load("/root/reference/checkpoints/birdnet_stm32n6_100_model_config.json").to_dict()
    default = config.ModelConfig().to_dict()
    json.dump({"shipped": shipped, "default": default}, open(os.path.join(HERE, "config_reference.json"), "w"), indent=1)

    # ---- oracle self-check ---------------------------------------------------------------------------
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "birdnet-stm32_b200"))
    from birdnet_stm32.audio import synth
    from birdnet_stm32.conversion.export_blob import export_blob
    from oracle import bn_oracle

    fx = os.path.join(ROOT, "tests", "fixtures")
    cfg = json.load(open(os.path.join(fx, "birdnet_stm32n6_100_model_config.json")))
    blob = export_blob(os.path.join(fx, "birdnet_stm32n6_100.tflite"), cfg)
    pcm = synth.synth_pcm16(8, 66150, 22050, seed=1234, edge_cases=True)
    peak = synth.file_peaks(pcm)
    spec = bn_oracle.frontend_hybrid(pcm, peak, 512, 66150 // 256, 256)
    model = bn_oracle.OracleModel(blob)
    scores, t96 = model.run(spec, tap_id=96)
    _, t127 = model.run(spec, tap_id=127)
    np.savez_compressed(os.path.join(HERE, "oracle_selfcheck.npz"), scores=scores, t96=t96, t127=t127,
                        spec_sum=spec.astype(np.float64).sum(axis=(1, 2, 3)), spec_col=spec[:, :, 100, 0])
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
