"""Turns an `ncu --set full` report of `bench.py --profile-mode` into the two files kept under profiles/ (not product code):

    ncu -i gpurun_out/full.ncu-rep --page raw --csv > gpurun_out/full_raw.csv
    python tools/ncu_summarize.py gpurun_out/full_raw.csv profiles/r01f "<source note>"

writes  <prefix>_ncu_full_step_b256_summary.csv  (one row per captured launch of the LAST complete step)
and     <prefix>_ncu_traffic.json                (DRAM bytes per launch by bench.py kernel tag -> roofline.traffic)
"""
import csv, json, re, sys

COLS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0,
        "nsecond": 1e-6, "second": 1e3}


def tag_of(name, seen):
    def nth(base):
        i = seen.get(base, 0)
        seen[base] = i + 1
        return f"{base}_{i}"
    if "conv0_fwd_tc_kernel" in name: return "conv0_fwd"
    if "conv0_bwd_tc_kernel" in name: return "conv0_wgrad"
    m = re.search(r"conv_tc_kernel<\(?(?:int\))?(\d+), \(?(?:int\))?(\d+)", name)
    if m:
        bn, epi = int(m.group(1)), int(m.group(2))
        if epi == 0: return "conv1_fwd" if bn == 128 else "conv2_fwd"
        return "conv1_dgrad" if bn == 64 else "conv2_dgrad"
    m = re.search(r"wgrad_tc_kernel<\(?(?:int\))?(\d+)", name)
    if m: return "conv1_wgrad" if int(m.group(1)) == 64 else "conv2_wgrad"
    if "attention_fwd_stream_kernel" in name: return "vqa_attention_fwd"
    if "attention_bwd_stream_kernel" in name: return "vqa_attention_bwd"
    if "adam_multi_kernel" in name: return "vqa_adam_multi"
    if "lstm_persistent_fwd_kernel" in name: return "lstm_recurrence_fwd"
    if "lstm_persistent_bwd_kernel" in name: return "lstm_bwd_persistent"
    if "dropnorm_bwd_unpool_kernel" in name: return "dropnorm_bwd_unpool"
    if "dropnorm_fwd_kernel" in name: return "dropnorm_fwd"
    if "dropnorm_bwd_kernel" in name: return "dropnorm_bwd"
    if "gemm_tc_persistent_kernel" in name: return nth("gemm_persistent")
    if "unpool_bf16_kernel" in name: return nth("unpool")
    return nth(re.sub(r"[^A-Za-z0-9_]+", "_", name.split("(")[0])[-40:])


def main():
    raw, prefix = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else "ncu --set full --clock-control none, bench.py --profile-mode --batch 256"
    rows = list(csv.reader(open(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kn = col["Kernel Name"]

    def val(r, k):
        if k not in col or r[col[k]] in ("", "n/a"):
            return None
        x = float(r[col[k]].replace(",", ""))
        return x * UNIT.get(units[col[k]], 1.0)

    # the last complete step = launches after the second-to-last Adam up to and including the last Adam
    adam = [i for i, r in enumerate(body) if "adam_multi_kernel" in r[kn]]
    if len(adam) >= 2:
        body = body[adam[-2] + 1: adam[-1] + 1]
    elif len(adam) == 1:
        body = body[: adam[0] + 1]
    seen, traffic = {}, {}
    with open(prefix + "_ncu_full_step_b256_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["tag", "Kernel Name", "Grid Size", "Block Size"] + COLS)
        for r in body:
            t = tag_of(r[kn], seen)
            w.writerow([t, r[kn][:110], r[col["Grid Size"]], r[col["Block Size"]]] + [val(r, k) for k in COLS])
            traffic[t] = {"dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
                          "ncu_duration_ms": val(r, "gpu__time_duration.sum"),
                          "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")}
    json.dump({"source": note, "kernels": traffic}, open(prefix + "_ncu_traffic.json", "w"), indent=1)
    for t, k in traffic.items():
        gb = ((k["dram_bytes_read"] or 0) + (k["dram_bytes_write"] or 0)) / 1e9
        print(f"{t:24s} {k['ncu_duration_ms']:.4f} ms  {gb:.3f} GB  {gb / max(k['ncu_duration_ms'], 1e-9) :.2f} TB/s  tensor {k['tensor_pipe_active_pct']}")


if __name__ == "__main__":
    main()
