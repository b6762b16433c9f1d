"""manual_benchmark-compatible CLI (benchmarks/manual_benchmark.py, SURVEY.md
8(f2)): flag parsing and CSV schema on CPU; on the GPU the whole pipeline with
every stage checked against the CPU oracle (the reference's --check_result,
benchmarks/manual_benchmark.cu:277-287,368-390,487-512)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
import manual_benchmark as mb  # noqa: E402


def test_absl_style_flags_and_defaults():
    f = mb.parse_flags([])
    # defaults of benchmarks/manual_benchmark.cu:44-81
    assert (f["num_categories"], f["embed_width"], f["batch_size"], f["hotness"]) == \
        (1048576, 128, 1024, 1)
    assert f["compressed_grad"] and f["skip_grad_init"] and f["clear_caches"]
    assert not f["half_embedding_type"] and not f["enable_csv"]
    f = mb.parse_flags("--num_categories 100 --alpha=1.15 --half_embedding_type "
                       "--nocompressed_grad --enable_csv=false --hotness=4 "
                       "--csr_input true".split())
    assert f["num_categories"] == 100 and abs(f["alpha"] - 1.15) < 1e-9
    assert f["half_embedding_type"] and not f["compressed_grad"] and not f["enable_csv"]
    assert f["hotness"] == 4 and f["csr_input"]
    with pytest.raises(SystemExit):
        mb.parse_flags(["--no_such_flag"])


def test_csv_schema_matches_the_reference():
    assert mb.CSV_HEADER == ("num_categories,batch_size,hotness,alpha,embed_width,combine_mode,"
                             "is_csr,is_weighted,compressed_grad,skip_grad_init,name,"
                             "iterations,elapsed_time_ms,avg_time_ms,algo_bw_l2,algo_bw_dram")
    f = mb.parse_flags("--num_categories 1000000 --batch_size 32768 --hotness 16 --alpha=1.05 "
                       "--embed_width 32".split())
    line = mb.csv_line(f, "backward", 1000, 123.456, 2000.0, 300.5)
    assert line == "1000000,32768,16,1.05,32,kSum,0,0,1,1,backward,1000 ,123.46 ,0.12 ,2000.00,300.50"
    assert len(line.split(",")) == len(mb.CSV_HEADER.split(","))


CASES = [
    "--num_categories 20000 --embed_width 32 --batch_size 512 --hotness 8",
    "--num_categories 30000 --embed_width 128 --batch_size 256 --hotness 16 --alpha=1.15 "
    "--half_embedding_type --use_int64_indices",
    "--num_categories 5000 --embed_width 64 --batch_size 300 --hotness 12 --csr_input "
    "--weighted_sum --nocompressed_grad --noskip_grad_init",
    "--num_categories 9000 --embed_width 16 --batch_size 128 --hotness 5 --combine_mode concat",
    "--num_categories 9000 --embed_width 64 --batch_size 128 --hotness 9 --csr_input "
    "--combine_mode mean --bf16",
]


@pytest.mark.gpu
@pytest.mark.parametrize("flags", CASES)
def test_pipeline_matches_oracle_stage_by_stage(cuda_lib, oracle, flags, tmp_path):
    import torch
    import gpu_helpers as gh
    csv = tmp_path / "out.csv"
    f = mb.parse_flags((flags + f" --iterations 2 --enable_csv --csv_file {csv} "
                                "--noenable_stderr").split())
    st = {}

    def check(stage, t):
        if stage == "forward":
            st["h"] = {k: gh.to_host(v) if isinstance(v, torch.Tensor) else v
                       for k, v in t.items()}
            h = st["h"]
            want = oracle.forward(h["table"], h["indices"], h["offsets"], h["weights"],
                                  h["batch"], h["num_hots"], h["mode"], embed_width=h["width"])
            import helpers
            assert helpers.bits_equal(h["out"], want)
        elif stage == "transpose":
            h = st["h"]
            idt = h["indices"].dtype
            if f["combine_mode"] == "concat":
                rows = oracle.extract_row_ids_concat(h["indices"].shape[0], idt)
            elif f["csr_input"]:
                rows = oracle.extract_row_ids_csr(h["offsets"], h["batch"], idt)
            else:
                rows = oracle.extract_row_ids_fixed(h["batch"], f["hotness"], idt)
            c_idx, c_sid, c_w = oracle.transpose(rows, h["indices"], h["weights"])
            assert np.array_equal(t["t_idx"].cpu().numpy(), c_idx)
            assert np.array_equal(t["t_sid"].cpu().numpy(), c_sid)
            c_rem = None
            if t["remapped"] is not None:
                c_rem = oracle.compressed_grad_indices(c_idx)
                assert np.array_equal(t["remapped"].cpu().numpy(), c_rem)
            st["coo"] = (c_idx, c_sid, c_w, c_rem)
        else:
            import helpers
            h = st["h"]
            c_idx, c_sid, c_w, c_rem = st["coo"]
            c_grad, c_inv = oracle.backward(h["grad_y"], h["width"], t["grad_rows"], c_idx,
                                            c_sid, c_rem, c_w, acc_f32=True)
            assert helpers.value_equal(gh.to_host(t["grad"]), c_grad)
            if t["inv"] is not None:
                assert np.array_equal(t["inv"].cpu().numpy(), c_inv)

    assert mb.run(f, check=check) == 0
    lines = csv.read_text().strip().splitlines()
    assert lines[0] == mb.CSV_HEADER
    assert [ln.split(",")[10] for ln in lines[1:]] == ["forward", "transpose", "backward"]
