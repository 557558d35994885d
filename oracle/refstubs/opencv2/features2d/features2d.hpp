// TEST INFRASTRUCTURE ONLY: forwards to the OpenCV stand-in of oracle/refstubs (see refcv.h).
#include "refcv.h"
