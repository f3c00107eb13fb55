#!/usr/bin/env python
"""gpurun_out/r02_sweep.json (tools/sweep.py) [+ the 40 GiB bench line] -> profiles/r02_sweep.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
doc = json.load(open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_sweep.json")))
big = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "r02_bench_40gib.json")
out = os.path.join(ROOT, "profiles", "r02_sweep.md")
ix = doc["index"]
with open(out, "w") as f:
    f.write("# r02: read-length sweep (BASELINE configs[4]), ONT-like reads (configs[3], single-GPU form), and the 40 GiB point\n\n"
            f"One B200, synthetic index of {ix['index_gib']} GiB ({ix['n_kmers']:,} k-mers, {ix['species']} species), loaded once; every case = "
            f"{doc['gbp_per_case']} Gbp of synthetic reads (70 % drawn from the indexed genomes, 1 % substitutions up to 500 bp, 5 % above), "
            "resident in HBM, `mbl_classify_resident` timed over 2 steps after 1 warm-up (tools/sweep.py).  seq-mode 1 up to 500 bp, 3 above.  "
            "`merge GB/s` = algorithmic bytes of the merge kernel (S_diff + 4K + 16 Nq + 24 Nm per launch) over its CUDA-event time; the index is "
            "streamed once per sub-batch, so long reads (more k-mer slots per batch than the HBM budget holds at once) pay it several times.\n\n")
    f.write("| read length | reads | ms / step | reads/s | Gbp/s | sub-batches | matches / step | merge ms | merge GB/s | frac of HBM peak | extract | sort | merge | match sort | score |\n"
            "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    rows = doc["sweep"] + ([doc["ont"]] if doc.get("ont") else [])
    for r in rows:
        st = r["stages_ms"]
        f.write(f"| {r['read_len']} | {r['reads']:,} | {r['ms_per_step']:.1f} | {r['reads_per_s']:,.0f} | {r['gbp_per_s']:.2f} | {r['sub_batches']} | "
                f"{r['matches_per_step']:,.0f} | {r['merge_ms_per_step']:.1f} | {r['merge_gbs']:.0f} | {r['merge_frac_of_peak']:.3f} | "
                f"{st.get('extract', 0)} | {st.get('sort', 0)} | {st.get('merge', 0)} | {st.get('match_sort', 0)} | {st.get('score', 0)} |\n")
    if doc.get("ont"):
        f.write(f"\nONT-like row: 1 M reads, {doc['ont']['bases'] / 1e9:.2f} Gbp per step (generated in {doc['ont'].get('reads_gen_s')} s), "
                f"{doc['ont']['classified']:,} classified.\n")
    if os.path.exists(big):
        try:
            d = json.loads(open(big).read())
            c = d["config"]
            f.write("\n## The north-star operating point: 10 M x 150 bp SE reads against a 40 GiB-class index on ONE B200\n\n"
                    f"`python bench.py --db-gib 40 --steps 2 --warmup 1 --no-cpu-baseline`: index {c['index_gib']} GiB ({c['n_kmers']:,} k-mers, "
                    f"{c['db_n_tiles']:,} tiles, {c['db_n_jumbo']} jumbo; generated in {c['db_generator_passes']} value-range passes, {c['db_gen_s']} s; loaded in {c['db_load_s']} s).\n\n"
                    f"* device pipeline: **{d['value'] / 1e6:.2f} M reads/s = {d['value'] * 60 / 1e6:.0f} M reads/min** ({d['ms_per_step']:.1f} ms per 10 M reads; target >= 50 M reads/min); "
                    f"end to end from pinned host buffers: {d['e2e']['value'] / 1e6:.2f} M reads/s\n"
                    f"* {c['merge_queries_per_step']:,} of {c['query_kmers_per_step']:,} metamers pass the presence filter (a 5x larger index collides with 5x more random 8-residue words), "
                    f"{c['matches_per_step']:,} matches, {c['classified_per_step']:,} reads classified, {c['sub_batches']} sub-batches per step (the index is streamed once per sub-batch)\n"
                    f"* stages (ms per step): " + ", ".join(f"{k[3:]} {v:.1f}" for k, v in d["stages_ms_per_step"].items() if v) + "\n"
                    f"* merge kernel: {d['roofline']['achieved']:.0f} GB/s of algorithmic bytes = {d['roofline']['frac']:.3f} of the measured HBM peak, {d['roofline']['ms_per_launch']:.1f} ms per launch\n")
        except Exception as e:
            f.write(f"\n(40 GiB line unreadable: {e})\n")
print("wrote", out)
