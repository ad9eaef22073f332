"""gpurun_out/c5_sweep.json -> profiles/<tag>_c5_sweep.md"""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
j = json.load(open("gpurun_out/c5_sweep.json"))
with open(f"profiles/{tag}_c5_sweep.md", "w") as f:
    f.write(f"# BASELINE configs[4] on one B200: beam sweep with histogram pruning ({tag})\n\n"
            f"`python tools/c5_sweep.py`: 64k-word trigram-shaped network, {j['states']} states / {j['arcs']} arcs, 6000 tied 16-mix GMMs; "
            f"{j['results'][0]['utterances']} utterances ({j['frames']} frames) decoded {j['results'][0]['lanes']} at a time, features resident in HBM, "
            "CUDA events on the decoder's stream after one warm-up pass.  Work counters are the reference's "
            "(`WFSTDecoderLite.cpp:231-241`), per frame and utterance.\n\n"
            "| main beam | maxHyps | frames/s | xRT | active models | emitting hyps | end hyps | planted sequence recovered | capacity failures |\n"
            "|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    for r in j["results"]:
        f.write(f"| {r['main_beam']:.0f} | {r['max_hyps']} | {r['frames_per_s']:.0f} | {r['frames_per_s'] / 100:.0f} | "
                f"{r['active_models_per_frame']:.0f} | {r['emit_hyps_per_frame']:.0f} | {r['end_hyps_per_frame']:.0f} | "
                f"{r['planted_sequence_recovered']}/{r['utterances']} | {r['capacity_failures']} |\n")
print(open(f"profiles/{tag}_c5_sweep.md").read())
