// TEST INFRASTRUCTURE: forwards the MFC header name to the oracle stub.
#pragma once
#include "mfc_stub.h"
