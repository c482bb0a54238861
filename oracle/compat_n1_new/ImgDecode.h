// TEST INFRASTRUCTURE (oracle/compat_n1_new): in the `n1_new` build the reference's CjfifDecode includes "ImgDecode.h" and
// gets THIS repository's CimgDecode (jpegsnoop_b200/csrc/host/ImgDecode.h) instead of the reference's — the drop-in.
#pragma once
#ifndef JSGPU_HOST_EXTERNAL_TYPES
#define JSGPU_HOST_EXTERNAL_TYPES 1
#endif
#include "host/ImgDecode.h"          // -I../jpegsnoop_b200/csrc
