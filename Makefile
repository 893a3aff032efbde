# Convenience targets; the driver uses __graft_entry__.build() / smoke() and bench.py directly.
.PHONY: all lib app test gpu-test bench clean
all: lib app
lib:
	$(MAKE) -C eigenkernel_b200/csrc -j8
app: lib
	$(MAKE) -C app -j8
test: all
	python -m pytest tests -x -q -m "not gpu"
gpu-test: all
	python -m pytest tests -x -q -m gpu
bench: all
	python bench.py
clean:
	$(MAKE) -C eigenkernel_b200/csrc clean
	$(MAKE) -C app clean
