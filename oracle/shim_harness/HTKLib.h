/* TEST INFRASTRUCTURE ONLY.  HAVE_HTKLIB is not defined: nothing of HTK's own library is used. */
