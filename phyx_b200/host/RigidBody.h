// Source compatibility with the reference include layout: everything lives in phyx_host.h.
#pragma once
#include "phyx_host.h"
