# oracle/_ref/pcb_reader: the REFERENCE's own stand-alone C reader of the .pcb layout, compiled from
# the source where it lies under /root/reference (never copied into this repo).  It parses a v1
# ChebyshevApproximation file and evaluates one point; tests use it to check that the files this
# package writes are readable by reference code.  Note: it rebuilds nodes with
# cos() and matches nodes by exact equality, so its values agree with the Python reference only
# to ~1e-13, not bit for bit -- it validates the FORMAT, it is not the numerical oracle.
REF_SRC ?= /root/reference/examples/binary_reader/reader.c
CC      ?= gcc

_ref/pcb_reader: $(REF_SRC)
	mkdir -p _ref
	$(CC) -O2 -std=c99 -o $@ $(REF_SRC) -lm
