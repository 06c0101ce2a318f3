/* l12_tables.h -- constant tables of the Layer I / II path (minimp3.d:284-435), shared by the host prepass and the device
 * parser.  tests/test_reference_tables.py compares them with the reference's literals. */
#ifndef L12_TABLES_H
#define L12_TABLES_H
#include <stdint.h>

/* L12_subband_alloc_t (minimp3.d:172-175): rows of {tab_offset, code_tab_width, band_count} */
static const uint8_t L12_ALLOC_L1[1][3] = {{76, 4, 32}};
static const uint8_t L12_ALLOC_L2M2[3][3] = {{60, 4, 4}, {44, 3, 7}, {44, 2, 19}};
static const uint8_t L12_ALLOC_L2M1[4][3] = {{0, 4, 3}, {16, 4, 8}, {32, 3, 12}, {40, 2, 7}};
static const uint8_t L12_ALLOC_L2M1_LOWRATE[2][3] = {{44, 4, 2}, {44, 3, 10}};

/* g_bitalloc_code_tab (minimp3.d:389-398) */
static const uint8_t L12_BITALLOC_CODE_TAB[92] = {
    0, 17, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16,
    0, 17, 18, 3, 19, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16,
    0, 17, 18, 3, 19, 4, 5, 16,
    0, 17, 18, 16,
    0, 17, 18, 19, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
    0, 17, 18, 3, 19, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14,
    0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

/* g_deq_L12 (minimp3.d:354-365).  The D literals are doubles that the static initialiser converts to float; written
 * the same way here (no `f` suffix) so that both go decimal -> double -> float. */
static const float L12_DEQ[54] = {
    3.17891e-07, 2.52311e-07, 2.00259e-07, 1.36239e-07, 1.08133e-07, 8.58253e-08,
    6.35783e-08, 5.04621e-08, 4.00518e-08, 3.07637e-08, 2.44172e-08, 1.93799e-08,
    1.51377e-08, 1.20148e-08, 9.53615e-09, 7.50925e-09, 5.96009e-09, 4.73053e-09,
    3.7399e-09, 2.96836e-09, 2.35599e-09, 1.86629e-09, 1.48128e-09, 1.17569e-09,
    9.32233e-10, 7.39914e-10, 5.8727e-10, 4.65889e-10, 3.69776e-10, 2.93492e-10,
    2.32888e-10, 1.84843e-10, 1.4671e-10, 1.1643e-10, 9.24102e-11, 7.3346e-11,
    5.82112e-11, 4.62023e-11, 3.66708e-11, 2.91047e-11, 2.31004e-11, 1.83348e-11,
    1.45521e-11, 1.155e-11, 9.16727e-12, 3.17891e-07, 2.52311e-07, 2.00259e-07,
    1.90735e-07, 1.51386e-07, 1.20155e-07, 1.05964e-07, 8.41035e-08, 6.6753e-08};

#endif
