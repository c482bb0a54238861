// TEST INFRASTRUCTURE: empty stand-in (oracle/compat_n1, N1 build of the reference CjfifDecode)
#pragma once
#include "mfc_stub.h"
