/* A C99 consumer of include/qsv.h: what a cgo / Rust-bindgen / ctypes binding sees.  No GPU needed: struct layout, error
 * paths and the plan API (host side of the library).  Built and run by tests/test_abi.py. */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "qsv.h"

#define CHECK(c) do { if (!(c)) { printf("FAILED line %d: %s\n", __LINE__, #c); ++failed; } } while (0)

int main(void) {
    int failed = 0;
    CHECK(sizeof(qsv_op) == 56 && offsetof(qsv_op, controls) == 16 && offsetof(qsv_op, param) == 24 && offsetof(qsv_op, matrix) == 40);
    CHECK(sizeof(qsv_stats) == 72);
    CHECK(QSV_OK == 0 && QSV_GATE_CUSTOM == 24);

    /* QFT-5 as the reference writes it (tests/qft.rs:55-62): H(pos), then CRk(k, pos + k - 1) */
    qsv_op ops[15];
    uint32_t ctrl[15];
    size_t n_ops = 0;
    memset(ops, 0, sizeof ops);
    for (uint32_t pos = 0; pos < 5; ++pos) {
        ops[n_ops].kind = QSV_GATE_H;
        ops[n_ops].target = pos;
        ++n_ops;
        for (uint32_t k = 2; k <= 5 - pos; ++k) {
            ctrl[n_ops] = pos + k - 1;
            ops[n_ops].kind = QSV_GATE_CRK;
            ops[n_ops].target = pos;
            ops[n_ops].n_controls = 1;
            ops[n_ops].controls = &ctrl[n_ops];
            ops[n_ops].iparam = (int32_t)k;
            ++n_ops;
        }
    }
    CHECK(n_ops == 15);
    qsv_plan* plan = NULL;
    CHECK(qsv_plan_create(&plan, 5, 5, ops, n_ops, 0, 0, 1) == QSV_OK && plan != NULL);
    qsv_stats st;
    memset(&st, 0, sizeof st);
    CHECK(qsv_plan_stats(plan, &st) == QSV_OK && st.n_gates == 15 && st.n_passes >= 1);
    size_t n_steps = 0;
    CHECK(qsv_plan_num_steps(plan, &n_steps) == QSV_OK && n_steps == st.n_passes);
    uint8_t layout[5];
    CHECK(qsv_plan_get_layout(plan, 1, layout, sizeof layout) == QSV_OK && layout[0] == 0 && layout[4] == 4);
    CHECK(qsv_plan_destroy(plan) == QSV_OK);

    /* error behaviour: a control equal to the target is refused with a message */
    ctrl[1] = 0;
    plan = NULL;
    CHECK(qsv_plan_create(&plan, 5, 5, ops, n_ops, 0, 0, 1) == QSV_ERR_INVALID_ARG && plan == NULL);
    CHECK(strlen(qsv_plan_last_error()) > 0);
    CHECK(qsv_init_basis(NULL, 0) == QSV_ERR_INVALID_ARG);
    printf("%d failed\n", failed);
    return failed ? 1 : 0;
}
