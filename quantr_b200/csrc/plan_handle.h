// plan_handle.h — the opaque qsv_plan of include/qsv.h.
#pragma once
#include "plan.h"

struct qsv_plan {
    qsv::Plan plan;
    // set by the CUDA side once the schedule has been uploaded; frees the device copy
    void (*release_device)(qsv_plan*) = nullptr;
};
