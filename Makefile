# Convenience targets; the driver uses __graft_entry__.build() / pytest / bench.py directly.
PY ?= python
build:
	$(PY) -c "import __graft_entry__ as g; g.build()"
test-cpu: build
	$(PY) -m pytest tests -x -q -m "not gpu"
test-gpu: build
	$(PY) -m pytest tests -x -q -m gpu
bench: build
	$(PY) bench.py
golden:
	$(PY) tests/golden/make_golden.py
clean:
	$(MAKE) -C bloomsearch_b200/csrc clean; $(MAKE) -C oracle clean; $(MAKE) -C synth clean; $(MAKE) -C tools clean
.PHONY: build test-cpu test-gpu bench golden clean
