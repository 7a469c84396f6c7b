// oracle/ref_shim: stands in for <colmap/util/misc.h> (TEST INFRASTRUCTURE; nothing of it is used on the path).
#pragma once
