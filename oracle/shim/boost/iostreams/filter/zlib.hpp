// no compression filters in this build (HAVE_LIBZ undefined)
