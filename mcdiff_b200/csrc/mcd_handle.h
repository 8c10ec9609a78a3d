// The library handle (one per device), shared by the translation units that define C entry points.
#pragma once

struct mcd_handle {
  int device;
  int sm_count;
  int max_smem_optin;
  long long launches;
};
