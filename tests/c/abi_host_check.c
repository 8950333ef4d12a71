/* Plain C99 caller of the C ABI (include/crbm_b200.h), no Python and no torch: the host-side entry points of the ingest
 * (secomo/sequences.py:20-52,101-117 restated as crbm_fasta_pack_block_host / crbm_fasta_parse_block_host / crbm_pack_onehot_host)
 * are called on small inputs and checked against a naive packing written here. Built and run by tests/test_host_cpu.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "crbm_b200.h"

#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "FAILED line %d: %s (last error: %s)\n", __LINE__, #cond, crbm_last_error()); return 1; } } while (0)

static int code_of(char ch) {
    switch (ch) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; }
    return -1;
}

/* the row layout of include/crbm_b200.h: one zero word, then 16 bases per little-endian word, zero padding */
static void naive_row(const char *seq, int L, uint32_t *row, int row_words) {
    int i;
    memset(row, 0, sizeof(uint32_t) * (size_t)row_words);
    for (i = 0; i < L; ++i) row[1 + i / 16] |= (uint32_t)code_of(seq[i]) << (2 * (i % 16));
}

int main(void) {
    const char *fasta = "; comment before the first header\n>r1 first\nACGTACGTAC\nGTACGTACGT\nAC\n>r2\nacgtNcgtacgtacgtacgtac\n>r3 empty\n>r4\r\nTTTTTTTTTTGGGGGGGGGGCC\r\n";
    const char *want[2] = {"ACGTACGTACGTACGTACGTAC", "TTTTTTTTTTGGGGGGGGGGCC"};
    const int L = 22;
    int rw, Lio = 0, bad = 0, r, w, threads;
    long long nrec = 0, nskip = 0, nbad = 0;
    uint32_t rows[8 * 64], ref[64];
    unsigned char codes[8 * 64];
    float onehot[2 * 4 * 22];

    CHECK(crbm_abi_version() == 1);
    rw = crbm_packed_row_words(L);
    CHECK(rw >= 1 + (L + 15) / 16 && rw % 4 == 0 && rw <= 64);

    /* sizing call: no buffer yet, record length unknown */
    CHECK(crbm_fasta_pack_block_host((const unsigned char *)fasta, strlen(fasta), NULL, 0, 0, &Lio, &nrec, &nskip, &bad, 1) == -13);
    CHECK(Lio == L && nrec == 4);                                   /* four headers: an upper bound on the rows */
    for (threads = 0; threads <= 3; ++threads) {
        memset(rows, 0xff, sizeof rows);
        CHECK(crbm_fasta_pack_block_host((const unsigned char *)fasta, strlen(fasta), rows, rw, 8, &Lio, &nrec, &nskip, &bad, threads) == CRBM_OK);
        CHECK(nrec == 2 && nskip == 1 && Lio == L);
        for (r = 0; r < 2; ++r) {
            naive_row(want[r], L, ref, rw);
            for (w = 0; w < rw; ++w) CHECK(rows[r * rw + w] == ref[w]);
        }
    }
    /* the byte-code variant of the same pass agrees */
    Lio = 0;
    CHECK(crbm_fasta_parse_block_host((const unsigned char *)fasta, strlen(fasta), codes, 8, &Lio, &nrec, &nskip, &bad) == CRBM_OK);
    CHECK(nrec == 2 && nskip == 1 && Lio == L);
    for (r = 0; r < 2; ++r) for (w = 0; w < L; ++w) CHECK(codes[r * L + w] == (unsigned char)code_of(want[r][w]));

    /* errors of the reference reader: ragged records (-11), a letter outside ACGTN (-12, reported) */
    Lio = 0;
    CHECK(crbm_fasta_pack_block_host((const unsigned char *)">a\nACGT\n>b\nACG\n", 15, rows, crbm_packed_row_words(4), 8, &Lio, &nrec, &nskip, &bad, 1) == -11);
    Lio = 0;
    CHECK(crbm_fasta_pack_block_host((const unsigned char *)">a\nACGT\n>b\nACXT\n", 16, rows, crbm_packed_row_words(4), 8, &Lio, &nrec, &nskip, &bad, 1) == -12);
    CHECK(bad == 'X');
    CHECK(crbm_fasta_pack_block_host(NULL, 0, rows, rw, 8, &Lio, &nrec, &nskip, &bad, 1) == CRBM_ERR_ARG);
    CHECK(strlen(crbm_last_error()) > 0);

    /* one-hot float32 (N,1,4,L) -> the same rows; a position with two ones and one with none are counted */
    memset(onehot, 0, sizeof onehot);
    for (r = 0; r < 2; ++r) for (w = 0; w < L; ++w) onehot[(r * 4 + code_of(want[r][w])) * L + w] = 1.f;
    memset(rows, 0xff, sizeof rows);
    CHECK(crbm_pack_onehot_host(onehot, 2, L, 4 * L, L, rows, rw, &nbad, 2) == CRBM_OK);
    CHECK(nbad == 0);
    for (r = 0; r < 2; ++r) {
        naive_row(want[r], L, ref, rw);
        for (w = 0; w < rw; ++w) CHECK(rows[r * rw + w] == ref[w]);
    }
    onehot[(0 * 4 + 3) * L + 0] = 1.f;                               /* A and T at position 0 of record 0 */
    onehot[(1 * 4 + code_of(want[1][21])) * L + 21] = 0.f;           /* nothing at the last position of record 1 */
    CHECK(crbm_pack_onehot_host(onehot, 2, L, 4 * L, L, rows, rw, &nbad, 1) == CRBM_OK);
    CHECK(nbad == 2);
    CHECK(crbm_pack_onehot_host(onehot, 2, L, 4 * L, L, rows, rw - 4, &nbad, 1) == CRBM_ERR_ARG);
    printf("abi_host_check ok\n");
    return 0;
}
