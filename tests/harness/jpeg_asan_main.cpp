// AddressSanitizer harness of the host JPEG decoder (tests/test_jpeg_malformed_cpu.py): every file named on the command
// line is pushed through the three read entry points of the C-ABI; the process only has to survive (ASan aborts on any
// out-of-bounds access).  Prints one "<rc_info> <rc_read> <rc_batch>" line per file.
#include <cstdint>
#include <cstdio>
#include <initializer_list>
#include <vector>

#include "../../include/rgbnm_b200.h"

int main(int argc, char** argv) {
    for (int a = 1; a < argc; ++a) {
        std::FILE* f = std::fopen(argv[a], "rb");
        if (!f) return 2;
        std::vector<uint8_t> buf;
        uint8_t tmp[4096];
        size_t n;
        while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
        std::fclose(f);
        // exact-size heap copy: a one-byte over-read of the input is an ASan error
        uint8_t* data = new uint8_t[buf.size() ? buf.size() : 1];
        for (size_t i = 0; i < buf.size(); ++i) data[i] = buf[i];
        rgbnm_jpeg_info info;
        const int r0 = rgbnm_jpeg_info_from_memory(data, buf.size(), &info);
        int r1 = -1, r2 = -1;
        {
            const size_t ycap = 64 * 64 * 64, ccap = 2 * 32 * 32 * 64;
            std::vector<int16_t> y(ycap), c(ccap), q(192);
            int32_t dims[6], flag = 0;
            r1 = rgbnm_jpeg_read_coefficients(data, buf.size(), y.data(), ycap, c.data(), ccap, q.data(), dims, &flag);
            const uint8_t* ptrs[1] = {data};
            const size_t sizes[1] = {buf.size()};
            uint8_t flags[1];
            int32_t status[1] = {0};
            r2 = rgbnm_jpeg_decode_batch(ptrs, sizes, 1, 64, 64, y.data(), c.data(), q.data(), flags, status, 1);
            if (r2 == 0) r2 = status[0];
            // plan-first variant: stop rows below, inside and beyond the image (and a negative one) must stay in bounds
            for (int32_t last : {-5, 0, 17, 1000}) {
                const int32_t rows[1] = {last};
                (void)rgbnm_jpeg_decode_batch_rows(ptrs, sizes, 1, 64, 64, y.data(), c.data(), q.data(), flags, status, 1, rows);
            }
        }
        std::printf("%d %d %d\n", r0, r1, r2);
        delete[] data;
    }
    return 0;
}
