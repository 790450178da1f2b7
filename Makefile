# parity: reference Makefile:8-26 (build / publish / test) + native build targets
.PHONY: build native test test-gpu sass clean
native:
	python -m vantage6_b200.ops.build
build: native
	python setup.py sdist bdist_wheel
test:
	python -m pytest tests -x -q -m "not gpu"
test-gpu:
	python -m pytest tests -x -q -m gpu
sass:
	python -c "from vantage6_b200.ops.build import dump_sass; print(dump_sass())"
clean:
	-rm -rf build dist vantage6_b200/ops/_build vantage6_b200/ops/_C*.so
