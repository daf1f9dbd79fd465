"""profiles/traffic.json from an ncu CSV of `bench.py --steps 2 --warmup 3 --no-cpu-baseline` (read here, no GPU).

usage: python profiles/make_traffic.py launches.csv [key]
The CSV must carry gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum, smsp__inst_executed.sum for
every launch.  One step = the kernels from the 4th k_heads launch (first timed step_many) up to the 5th; for the
ray-march modes, the 4th k_step_many launch alone."""
import collections, csv, json, os, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    path = sys.argv[1]
    key = sys.argv[2] if len(sys.argv) > 2 else "step_cached"
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    iid, ik, im, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        launches.setdefault(int(r[iid]), {"name": r[ik].split("(")[0].replace("void ", "")})[r[im]] = float(r[iv].replace(",", ""))
    seq = list(launches.values())
    if key == "step_cached":
        # a step of the cached path = the kernels after an L2-flush fill that is followed by k_heads, up to the next
        # fill; the 4th such step is the first timed one of `bench.py --steps 2 --warmup 3`
        fills = [i for i, l in enumerate(seq) if l["name"].startswith("at::vectorized_elementwise_kernel")
                 and i + 1 < len(seq) and seq[i + 1]["name"].startswith("k_heads")]
        start = fills[3]
        end = next(i for i in range(start + 1, len(seq)) if seq[i]["name"].startswith("at::"))
        step = seq[start + 1:end]
    else:
        idx = [i for i, l in enumerate(seq) if l["name"].startswith("k_step_many")]
        step = [seq[idx[3]]]
    step = [l for l in step if l["name"].startswith("k_")]
    tot_t = sum(l["gpu__time_duration.sum"] for l in step)
    share = collections.OrderedDict()
    for l in step:
        n = l["name"].split("<")[0]
        share[n] = share.get(n, 0.0) + l["gpu__time_duration.sum"] / tot_t
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    try:
        data = json.load(open(out_path))
    except Exception:
        data = {}
    data[key] = {
        "dram_bytes_per_launch": int(sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in step)),
        "warp_inst_per_launch": int(sum(l["smsp__inst_executed.sum"] for l in step)),
        "kernels_per_step": len(step),
        "serialized_us": round(tot_t / 1e3, 1),
        "share": {k: round(v, 3) for k, v in share.items()},
        "csrc": __import__("bench").csrc_stamp(),   # bench.py replays these figures only for the same kernel sources
        "source": os.path.basename(path) + " (ncu launch list, kernels serialised: use shares and counters, not the absolute time)",
    }
    json.dump(data, open(out_path, "w"), indent=1)
    print(json.dumps(data[key], indent=1))


if __name__ == "__main__":
    main()
