"""`import dct_manip` (datasets.py:10, utils/custom_transforms.py:7 of the reference) -> the B200 repository's host decoder.
Same return contract as dct_manip.cpp:152-178 / :578-606; see INTEGRATION.md section 1."""
from rgb_no_more_b200.dct_manip import (decode_batch, decode_coeff, quantize_at_quality, read_coefficients,  # noqa: F401
                                        read_coefficients_from_bytes, write_coefficients)
