// placeholder translation unit: the device N_Vector lands here
