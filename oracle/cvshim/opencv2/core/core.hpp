// TEST INFRASTRUCTURE ONLY: forwards to the cv shim (see cvshim.h).
#pragma once
#include "../../cvshim.h"
