"""Known-answer vectors transcribed from the reference's own gtest suites
(NVIDIA/cuEmbed @ 90dd8436; file:line relative to the reference root).
Used to pin the CPU oracle (tests/test_oracle.py) and the CUDA path
(tests/test_*_gpu.py)."""

# ---- tests/test_embedding_forward.cu:118-160 -------------------------------
FWD = dict(
    embed_width=4,
    hotness=2,
    batch_size=2,
    embedding=[float(i) for i in range(1, 21)],        # :121-123, 5 rows x 4
    indices=[1, 3, 0, 4],                              # :125
    offsets=[0, 2, 4],                                 # :126
    weights=[1.0, 0.5, 1.0, 0.5],                      # :127
    concat=[5., 6., 7., 8., 13., 14., 15., 16., 1., 2., 3., 4., 17., 18., 19., 20.],  # :128-129
    sum=[18., 20., 22., 24., 18., 20., 22., 24.],      # :130-139
    avg=[9., 10., 11., 12., 9., 10., 11., 12.],        # :140-149
    sum_weighted=[11.5, 13., 14.5, 16., 9.5, 11., 12.5, 14.],  # :150-159
)

# ---- tests/test_embedding_transpose.cu:112-122 ------------------------------
TRANSPOSE = dict(
    nnz=4,
    indices=[1, 3, 0, 4],
    sample_ids=[0, 0, 1, 1],
    weights=[1.0, 0.5, 1.0, 0.5],
    transpose_indices=[0, 1, 3, 4],
    transpose_sample_ids=[1, 0, 0, 1],
    transpose_sample_ids_concat=[2, 0, 1, 3],   # sample ids 0..3 (ExtractRowIdsForConcat)
    transpose_weights=[1.0, 1.0, 0.5, 0.5],
)

# ---- tests/test_embedding_backward.cu:162-202 -------------------------------
BWD = dict(
    embed_width=4,
    num_categories=5,
    batch_size=2,
    nnz=4,
    num_unique=3,
    transpose_indices=[0, 1, 3, 3],
    transpose_remapped_indices=[0, 1, 2, 2],
    transpose_sample_ids=[1, 0, 0, 1],
    transpose_sample_ids_concat=[2, 0, 1, 3],
    transpose_weights=[3.0, 1.0, 0.5, 3.0],
    grad_y_sum=[1., 2., 3., 4., 5., 6., 7., 8.],
    grad_y_concat=[float(i) for i in range(1, 17)],
    grad_sum=[5., 6., 7., 8., 1., 2., 3., 4., 0., 0., 0., 0., 6., 8., 10., 12., 0., 0., 0., 0.],
    grad_sum_weighted=[15., 18., 21., 24., 1., 2., 3., 4., 0., 0., 0., 0., 15.5, 19., 22.5, 26., 0., 0., 0., 0.],
    grad_concat=[9., 10., 11., 12., 1., 2., 3., 4., 0., 0., 0., 0., 18., 20., 22., 24., 0., 0., 0., 0.],
    grad_concat_weighted=[27., 30., 33., 36., 1., 2., 3., 4., 0., 0., 0., 0., 41.5, 45., 48.5, 52., 0., 0., 0., 0.],
    inverse_mapping=[0, 1, 3],
    cgrad_sum=[5., 6., 7., 8., 1., 2., 3., 4., 6., 8., 10., 12.],
    cgrad_sum_weighted=[15., 18., 21., 24., 1., 2., 3., 4., 15.5, 19., 22.5, 26.],
    cgrad_concat=[9., 10., 11., 12., 1., 2., 3., 4., 18., 20., 22., 24.],
    cgrad_concat_weighted=[27., 30., 33., 36., 1., 2., 3., 4., 41.5, 45., 48.5, 52.],
)

# ---- cuembed/README.md:132,141,150,202 (documentation examples) -------------
README = dict(
    fixed_num_hots=3, fixed_row_ids=[0, 0, 0, 1, 1, 1, 2, 2, 2],
    csr_offsets=[0, 2, 3, 5], csr_row_ids=[0, 0, 1, 2, 2],
    concat_row_ids=[0, 1, 2, 3],
    compress_in=[4, 4, 7, 8, 8, 8, 18], compress_out=[0, 0, 1, 2, 2, 2, 3],
)

# ---- tests/test_embedding_against_cpu.cu:236-293: the randomised shape matrix
# (batch, width, hotness, mode, csr, weighted, compressed), 20 K categories.
def against_cpu_matrix():
    out = []
    def add(bs, w, h, mode, csr=False, weighted=False, cmp=False):
        out.append(dict(batch=bs, width=w, hot=h, mode=mode, csr=csr,
                        weighted=weighted, compressed=cmp))
    for w in (2, 4):                                   # :237-254
        add(3, w, 4, "sum"); add(3, w, 4, "sum", csr=True)
        if w == 2:
            add(3, w, 4, "sum", weighted=True); add(3, w, 4, "sum", csr=True, weighted=True)
            add(3, w, 4, "mean"); add(3, w, 4, "mean", csr=True)
        else:
            add(3, w, 4, "mean"); add(3, w, 4, "mean", csr=True)
            add(3, w, 4, "sum", weighted=True); add(3, w, 4, "sum", csr=True, weighted=True)
        add(3, w, 4, "concat"); add(3, w, 4, "sum", cmp=True); add(3, w, 4, "concat", cmp=True)
    for w in (32, 36):                                 # :255-272
        add(1023, w, 26, "sum"); add(1023, w, 26, "sum", csr=True)
        add(1023, w, 26, "sum", weighted=True); add(1023, w, 26, "sum", csr=True, weighted=True)
        add(1023, w, 26, "mean"); add(1023, w, 26, "mean", csr=True)
        add(1023, w, 26, "concat"); add(1023, w, 26, "sum", cmp=True); add(1023, w, 26, "concat", cmp=True)
    # :273-281
    add(3, 512, 63, "sum"); add(3, 512, 63, "sum", csr=True)
    add(3, 512, 63, "sum", weighted=True); add(3, 512, 63, "sum", csr=True, weighted=True)
    add(3, 512, 63, "mean"); add(3, 512, 63, "mean", csr=True)
    add(3, 512, 63, "concat"); add(3, 512, 63, "sum", cmp=True); add(3, 512, 63, "concat", cmp=True)
    # :282-286
    add(1023, 512, 63, "sum"); add(1023, 512, 63, "sum", csr=True)
    add(1023, 512, 63, "sum", weighted=True); add(1023, 512, 63, "sum", csr=True, weighted=True)
    add(1023, 512, 63, "concat")
    # :287-293
    add(1023, 514, 63, "sum"); add(1023, 514, 63, "sum", csr=True)
    add(1023, 514, 63, "sum", weighted=True); add(1023, 514, 63, "sum", csr=True, weighted=True)
    add(1023, 514, 63, "concat"); add(1023, 514, 63, "sum", cmp=True); add(1023, 514, 63, "concat", cmp=True)
    return out
