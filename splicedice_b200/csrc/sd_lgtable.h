// sd_lgtable.h -- double-double log-factorial table (host builder; see sd_lgtable.cpp).
#pragma once
#include <stdint.h>

#include <vector>

namespace sd {
void lgtable_host(int64_t entries, std::vector<double> *out);
}
