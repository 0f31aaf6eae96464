// TEST INFRASTRUCTURE ONLY.  The reference includes "misc/utils.h" (color_balance.cpp:1) from the
// wider CUAUV tree but uses nothing from it; an empty header satisfies the include.
#pragma once
