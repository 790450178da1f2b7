# Targets of the reference Makefile (build / publish / test, Makefile:8-26) plus the native build of this repo.
PYTHON ?= python

.PHONY: help native build publish test test-gpu coverage sass clean

help:
	@echo "native    compile the sm_100a extension in-tree (nvcc, ~20 s)"
	@echo "build     native + sdist + wheel into dist/"
	@echo "publish   upload dist/* with twine (needs credentials)"
	@echo "test      CPU test-suite (pytest -m 'not gpu')"
	@echo "test-gpu  GPU test-suite (needs a B200)"
	@echo "coverage  CPU test-suite under coverage, if coverage is installed"
	@echo "sass      cuobjdump -sass summary of the hand-written kernels"

native:
	$(PYTHON) -m vantage6_b200.ops.build

build: native
	$(PYTHON) setup.py sdist bdist_wheel

publish: build
	$(PYTHON) -m twine upload --repository pypi dist/*

test:
	$(PYTHON) -m pytest tests -x -q -m "not gpu"

test-gpu:
	$(PYTHON) -m pytest tests -x -q -m gpu

coverage:
	$(PYTHON) -m coverage run --source=vantage6_b200 -m pytest tests -q -m "not gpu" && $(PYTHON) -m coverage report

sass:
	$(PYTHON) scripts/sass_report.py

clean:
	-rm -rf build dist *.egg-info vantage6_b200/ops/_build vantage6_b200/ops/_C*.so
