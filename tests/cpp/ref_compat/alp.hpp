// alp.hpp — what a program written against the reference's umbrella header (include/alp.hpp:1-15) finds when it is
// pointed at this repository instead: the same primitive API, served by libalp_b200.so (include/alp_b200.hpp).
#pragma once
#include "alp_b200.hpp"
