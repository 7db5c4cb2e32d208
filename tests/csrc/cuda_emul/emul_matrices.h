// TEST INFRASTRUCTURE -- CSR test matrices of the SpMV harnesses (complex values; the real harness takes the real parts)
#pragma once
#include <string.h>

#include <complex>
#include <random>
#include <vector>

typedef std::complex<double> CD;

struct Csr {
    long long nrows, ncols;
    std::vector<int> rowptr, colidx;
    std::vector<CD> vals;
};

static Csr make_matrix(const char* kind, std::mt19937_64& rng) {
    std::normal_distribution<double> nd;
    Csr A;
    auto finish_row = [&]() { A.rowptr.push_back((int)A.colidx.size()); };
    auto put = [&](int c) {
        A.colidx.push_back(c);
        A.vals.push_back(CD(nd(rng), nd(rng)));
    };
    A.rowptr.push_back(0);
    if (!strcmp(kind, "stencil5")) {                      // 5 per row: CPR 6; 2563 rows = 11 tiles
        const int nx = 11, ny = 233;
        A.nrows = A.ncols = (long long)nx * ny;
        for (int r = 0; r < nx * ny; ++r) {
            const int i = r / ny, j = r % ny;
            if (i > 0) put(r - ny);
            if (j > 0) put(r - 1);
            put(r);
            if (j < ny - 1) put(r + 1);
            if (i < nx - 1) put(r + ny);
            finish_row();
        }
    } else if (!strcmp(kind, "band7")) {                  // 7 per row: CPR 8
        A.nrows = A.ncols = 2100;
        const int offs[7] = {-30, -2, -1, 0, 1, 2, 30};
        for (int r = 0; r < 2100; ++r) {
            for (int o : offs)
                if (r + o >= 0 && r + o < 2100) put(r + o);
            finish_row();
        }
    } else if (!strcmp(kind, "rand12")) {                 // ~12 per row: CPR 16, ragged rows incl. empty ones
        A.nrows = 1500;
        A.ncols = 1700;
        for (int r = 0; r < 1500; ++r) {
            const int cnt = (r % 7 == 3) ? 0 : (int)(rng() % 25);
            int c = (int)(rng() % 40);
            for (int t = 0; t < cnt && c < 1700; ++t) {
                put(c);
                c += 1 + (int)(rng() % 60);
            }
            finish_row();
        }
    } else if (!strcmp(kind, "ragged")) {                 // mostly one entry per row, one tile far beyond the stage
        A.nrows = A.ncols = 3000;                         // capacity (direct global loads), nnz % 4 != 0
        for (int r = 0; r < 3000; ++r) {
            if (r >= 600 && r < 606) {
                for (int c = r % 2; c < 3000; c += 2) put(c);
            } else if (r % 11 != 5) {
                put(r);
            }
            finish_row();
        }
        if (A.colidx.size() % 4 == 0) {                   // force an unaligned tail
            A.colidx.push_back(2999);
            A.vals.push_back(CD(1.5, -0.5));
            A.rowptr.back() += 1;
        }
    } else if (!strcmp(kind, "tiny")) {
        A.nrows = A.ncols = 3;
        put(0); put(1); finish_row();
        finish_row();
        put(1); put(2); finish_row();
    } else if (!strcmp(kind, "long")) {                   // ~60 per row: warp-per-row kernel
        A.nrows = 300;
        A.ncols = 900;
        for (int r = 0; r < 300; ++r) {
            for (int c = (int)(rng() % 15); c < 900; c += 1 + (int)(rng() % 28)) put(c);
            finish_row();
        }
    } else {
        fprintf(stderr, "unknown matrix kind %s\n", kind);
        exit(2);
    }
    return A;
}

