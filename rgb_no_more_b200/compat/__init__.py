"""Drop-in seams for the reference's own `train.py` (boundaries B2 and B4).

    python -m rgb_no_more_b200.compat.launch --ref /path/to/RGB-no-more --backend b200 -- \
        --train --domain dct --model_arch vits --indexpaths train.csv,val.csv --temp_datapath /data --no_extract --no_resize

runs the REFERENCE's `train.py` (its CLI, config, loop, evaluation, checkpointing) on the target PyTorch with
  * `env.activate`  -- stand-ins for the third-party modules the reference imports but the image lacks (`shims/`), the
                       B1 `dct_manip` drop-in, and the `CosineAnnealingLR(verbose=)` keyword torch >= 2.7 rejects;
  * `backend.install` (only with --backend b200 / RGBNM_BACKEND=b200) -- `utils.get_model` returns the B200 model classes,
                       `utils.get_dataset` returns `loader.B200Loader`s (JPEG bytes -> host Huffman decode -> fused K0 on the GPU)
                       and `utils.unpack_data` applies RandomMixup_DCT on the device batch.
Without the opt-in the reference path runs untouched (`--backend reference`)."""
